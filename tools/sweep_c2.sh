#!/bin/bash
# development sweep of the C2 island step: islands per GPU x CTA shape (device-timed value only)
out=${1:-gpurun_out/sweep_c2.txt}
shift
: > $out
cfgs=("$@")
if [ ${#cfgs[@]} -eq 0 ]; then cfgs=("592 256 4" "888 256 4" "1184 256 4" "1776 256 4" "2368 256 4"); fi
for cfg in "${cfgs[@]}"; do
  set -- $cfg
  r=$(GJ_FUSED_THREADS=$2 GJ_FUSED_MB=$3 GJ_BENCH_EXTRAS=0 python bench.py --steps 20 --warmup 3 --no-cpu-baseline --no-e2e --islands $1 2>/dev/null | python -c "import sys,json; d=json.loads(sys.stdin.read()); print('%.3f G/s  step %.1f us  kernel %.1f us  %s' % (d['value']/1e9, d['ms_per_step']*1e3, d['roofline']['kernel_ms']*1e3, d['step_path']))")
  echo "islands=$1 threads=$2 mb=$3 : $r" | tee -a $out
done
