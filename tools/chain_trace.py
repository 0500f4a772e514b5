#!/usr/bin/env python
"""Per-phase timeline of one CTA of the persistent conv chain (ESRP_CHAIN_TRACE, csrc/esrp_conv_chain.cu): where a
phase's time goes.  Slots: 0 producer at phase start, 1 issuers done issuing the previous phase, 2 previous MMAs complete
(weights fetch starts), 3 neighbours' flags acquired, 4 last row load issued, 7/8 first row of issuer 0/1 about to issue, 9/10 issuer 0/1 done, 11 epilogue thread 0 done with its rows, 12 flag published."""
import json, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
path = "/tmp/chain_trace.txt"
os.environ["ESRP_CHAIN_TRACE"] = path
import torch
import esrganplus_b200 as E
from esrganplus_b200.synth import random_state_dict_g
nb = int(sys.argv[1]) if len(sys.argv) > 1 else 4
dev = torch.device("cuda:0")
net = E.RRDBNet(3, 3, 64, nb); net.load_state_dict(random_state_dict_g(3, 3, 64, nb, seed=31)); net = net.to(dev).eval()
for p in net.parameters(): p.requires_grad = False
x = torch.rand(16, 3, 128, 128, device=dev)
with torch.no_grad():
    for _ in range(3): net(x)
torch.cuda.synchronize()
rows = [[int(v) for v in l.split()] for l in open(path)]
clk = 1.0   # cycles; report in kilocycles
names = ["conv1", "conv2", "conv3", "conv4", "conv5"]
agg = {}
for i, r in enumerate(rows):
    if i == 0 or i + 1 >= len(rows): continue
    nxt = rows[i + 1]
    first_issue = min(v for v in (r[7], r[8]) if v)
    last_issue = max(r[9], r[10])
    d = {"drain_wait(0->2)": r[2] - r[0], "flag_wait(2->3)": r[3] - r[2], "w+first_row(3->first MMA)": first_issue - r[3],
         "issue(first->last MMA)": last_issue - first_issue, "epi_tail(last MMA->flag)": r[12] - last_issue,
         "phase(flag->flag)": r[12] - rows[i - 1][12], "prod_idle_end(4->next 0)": nxt[0] - r[4]}
    a = agg.setdefault(names[i % 5], {})
    for k, v in d.items(): a.setdefault(k, []).append(v)
out = {n: {k: round(sum(v) / len(v) / 1000.0, 2) for k, v in a.items()} for n, a in agg.items()}
print(json.dumps({"unit": "kilo-cycles (avg over blocks)", "cta": os.environ.get("ESRP_CHAIN_TRACE_CTA", "70"), "phases": out}, indent=1))
