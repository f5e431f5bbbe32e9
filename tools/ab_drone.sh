#!/bin/bash
# A/B over ab/lib_*.so on the drone config (step only, random controls): device-timed step, then DRAM bytes and executed
# local-memory instructions of k_step per variant (ncu, a few launches), then the drone GPU tests on the last variant.
mkdir -p gpurun_out
for round in 1 2; do
  for f in ab/lib_*.so; do
    cp "$f" mujoco-template_b200/libb2mj.so
    echo -n "$(basename $f) r$round: "
    python tools/bench_value.py --model drone --no-linearize --no-secondary --steps 300 2>&1 | tail -1
    echo -n "   cartpole step only: "
    python tools/bench_value.py --model cartpole --no-linearize --no-secondary --steps 300 2>&1 | tail -1 | cut -c1-60
  done
done
for f in ab/lib_*.so; do
  cp "$f" mujoco-template_b200/libb2mj.so
  n=$(basename $f .so)
  ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum,sass__inst_executed_local_loads,sass__inst_executed_local_stores,smsp__inst_executed.sum \
      --clock-control none -k regex:k_step -c 6 --csv --log-file gpurun_out/drone_$n.csv \
      python bench.py --model drone --no-linearize --no-secondary --no-cpu-baseline --no-e2e --no-graph --steps 3 --warmup 3 > /dev/null 2>&1
  echo "== $n"; python - "$n" <<'P'
import csv, sys, collections
rows=[r for r in csv.reader(open("gpurun_out/drone_%s.csv" % sys.argv[1])) if len(r)>10]
h=rows[0]; im,iv,iid=h.index("Metric Name"),h.index("Metric Value"),h.index("ID")
d=collections.defaultdict(dict)
for r in rows[1:]: d[r[iid]][r[im]]=r[iv]
for k,m in d.items(): print(k, m)
P
done
cp ab/lib_${FINAL:-latelazy}.so mujoco-template_b200/libb2mj.so
ncu --set full --clock-control none --import-source on -k regex:k_step -c 1 -f -o gpurun_out/r03l_drone_step python tools/prof_model.py drone 262144 > gpurun_out/r03l_ncu.log 2>&1
python -m pytest tests -q -m gpu 2>&1 | tail -4
