"""Host-side check of the persistent conv chain's dependency arithmetic (csrc/conv_chain_dep.h): compiled with g++ and
compared with a brute-force enumeration of which row groups own the rows a phase reads / overwrites."""
import os
import subprocess
import sys
import textwrap

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_chain_dep_range_covers_every_overlapping_group(tmp_path):
    src = tmp_path / "t.cpp"
    src.write_text(textwrap.dedent(r'''
        #include <cstdio>
        #include <vector>
        #include "conv_chain_dep.h"
        using namespace esrp;
        int main() {
          long long checked = 0;
          const long long Us[] = {1, 2, 3, 7, 74, 147, 148, 149, 296, 300, 1000, 2048, 4099};
          const int grids[] = {1, 2, 4, 6, 74, 148};
          for (long long U : Us) for (int grid : grids) for (int nsl : {1, 2}) for (int nslq : {1, 2}) {
            if (grid % nsl || grid % nslq) continue;
            const int ng = grid / nsl, ngq = grid / nslq;
            // owner of every unit in the previous phase's split, by enumeration
            std::vector<int> owner(U, -1);
            for (int g = 0; g < ngq; ++g)
              for (long long u = chain_group_start(U, g, ngq); u < chain_group_start(U, g + 1, ngq); ++u) owner[u] = g;
            for (long long u = 0; u < U; ++u) {
              if (owner[u] < 0) { printf("unit %lld of U=%lld ngq=%d has no owner\n", u, U, ngq); return 1; }
              if (chain_group_of(U, u, ngq) != owner[u]) { printf("group_of U=%lld u=%lld ngq=%d\n", U, u, ngq); return 1; }
            }
            for (int g = 0; g < ng; ++g) {
              int lo, hi;
              chain_dep_range(U, g, ng, ngq, &lo, &hi);
              if (lo < 0 || hi >= ngq || lo > hi) { printf("range out of bounds U=%lld g=%d\n", U, g); return 1; }
              long long a = chain_group_start(U, g, ng) - 1, b = chain_group_start(U, g + 1, ng);
              for (long long u = a; u <= b; ++u) {
                if (u < 0 || u >= U) continue;
                if (owner[u] < lo || owner[u] > hi) {
                  printf("U=%lld grid=%d nsl=%d nslq=%d g=%d: unit %lld owned by %d outside [%d,%d]\n", U, grid, nsl, nslq, g, u, owner[u], lo, hi);
                  return 1;
                }
                ++checked;
              }
            }
          }
          printf("ok %lld\n", checked);
          return 0;
        }
    '''))
    exe = tmp_path / "t"
    subprocess.run(["g++", "-O1", "-std=c++17", "-I", os.path.join(ROOT, "esrganplus_b200", "csrc"), str(src), "-o", str(exe)],
                   check=True)
    r = subprocess.run([str(exe)], capture_output=True, text=True)
    assert r.returncode == 0 and r.stdout.startswith("ok"), r.stdout + r.stderr
