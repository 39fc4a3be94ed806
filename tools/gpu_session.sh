#!/bin/bash
# One GPU session: tests, both bench arms, ncu launch list and full captures of the two kernels of the step.
# Usage (on the GPU box, from the repo root): bash tools/gpu_session.sh <tag>
TAG=${1:-r1}
OUT=gpurun_out
mkdir -p $OUT
python -m pytest tests -m gpu -x -q 2>&1 | tail -6 > $OUT/${TAG}_pytest_gpu.log
python __graft_entry__.py --smoke > $OUT/${TAG}_smoke.log 2>&1
python bench.py --impl reference --steps 10 --warmup 3 > $OUT/${TAG}_bench_reference.json 2> $OUT/${TAG}_bench_reference.err
python bench.py --steps 50 --warmup 5 > $OUT/${TAG}_bench.json 2> $OUT/${TAG}_bench.err
# launch list of the same command (cold-cache, serialised times: only the SHARES are comparable)
ncu --metrics gpu__time_duration.sum --clock-control none -c 40 --csv --log-file $OUT/${TAG}_launches.csv \
    python bench.py --steps 6 --warmup 3 > $OUT/${TAG}_ncu_launches.log 2>&1
# one full capture per kernel (steady state: skip the first launches)
# (perf_probe: one warm-up advance launch of 8 steps, then one advance launch of 12 steps with snapshots + 12 renders)
python tools/perf_probe.py --worlds 131072 --balls 8 --steps 12 --render --warm 8 --max-substeps 64 > $OUT/${TAG}_probe_plain.log 2>&1 && \
for KS in advance_kernel:1 render_tile_kernel:4; do
  K=${KS%%:*}; S=${KS##*:}
  ncu --set full --import-source on --clock-control none -k regex:$K -s $S -c 1 -f -o $OUT/${TAG}_full_$K \
      python tools/perf_probe.py --worlds 131072 --balls 8 --steps 12 --render --warm 8 --max-substeps 64 > $OUT/${TAG}_ncu_$K.log 2>&1
done
[ -x tools/_bin/store_ceiling ] && tools/_bin/store_ceiling > $OUT/${TAG}_store_ceiling.log 2>&1
python tools/sweep.py --out $OUT/${TAG}_sweep.json > $OUT/${TAG}_sweep.log 2>&1
python tools/steps_probe.py > $OUT/${TAG}_steps.log 2>&1
python tools/native_probe.py > $OUT/${TAG}_native.log 2>&1
python tools/e2e_probe.py > $OUT/${TAG}_e2e.log 2>&1
tail -3 $OUT/${TAG}_pytest_gpu.log
tail -2 $OUT/${TAG}_smoke.log
head -c 600 $OUT/${TAG}_bench.json; echo
head -c 400 $OUT/${TAG}_bench_reference.json; echo
