#!/bin/bash
# Round-2 profile refresh: ncu launch lists (config 2 forward, config-4-shaped train fwd+bwd of the generator, discriminator
# fwd+bwd), full capture of six row-kernel launches; bench twice (is the fp32 e2e leg's copy overlapped?).
mkdir -p gpurun_out
M=gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum
timeout 900 ncu --metrics $M --clock-control none --profile-from-start off --csv \
   --log-file gpurun_out/y_launches_c2.csv python tools/profile_step.py > gpurun_out/y_ncu_list_c2.log 2>&1; echo "ncu list c2 rc=$?"
timeout 900 ncu --metrics $M --clock-control none --profile-from-start off --csv \
   --log-file gpurun_out/y_launches_train_crop32.csv python tools/profile_step.py --batch 32 --tile 32 --bwd --train > gpurun_out/y_ncu_list_train.log 2>&1; echo "ncu list train rc=$?"
timeout 900 ncu --metrics $M --clock-control none --profile-from-start off --csv \
   --log-file gpurun_out/y_launches_d.csv python tools/profile_d.py > gpurun_out/y_ncu_list_d.log 2>&1; echo "ncu list d rc=$?"
timeout 900 ncu --set full --clock-control none --import-source on --profile-from-start off \
   -k regex:conv3x3_row -s 6 -c 6 -f -o gpurun_out/y_prof python tools/profile_step.py > gpurun_out/y_ncu_full.log 2>&1; echo "ncu full rc=$?"
ncu -i gpurun_out/y_prof.ncu-rep --page raw --csv > gpurun_out/y_prof_raw.csv 2>/dev/null; echo "raw rc=$?"
for i in 1 2; do
  timeout 900 python bench.py --no-extras > gpurun_out/y_bench_$i.json 2> gpurun_out/y_bench_err_$i.log; echo "bench $i rc=$?"
done
python - <<'PY'
import json
for i in (1, 2):
    d = json.loads(open(f"gpurun_out/y_bench_{i}.json").read().strip().splitlines()[-1])
    print(i, d["value"], d["ms_per_step"], d["e2e"], d["roofline"]["frac"], d["clocks"])
PY
