#!/bin/bash
# Profile evidence for one round (run under gpurun, 1 GPU):  bash tools/collect_profiles.sh r01
# 1. ncu launch list over ~one step of the exact bench command   2. ncu --set full of the roofline kernels
R=${1:-r01}
O=gpurun_out
mkdir -p $O
python bench.py --steps 1 --warmup 3 --no-cpu-baseline > $O/bench_plain_$R.log 2>&1
SKIP=$(grep -o "launches before the timed region: [0-9]*" $O/bench_plain_$R.log | grep -o "[0-9]*$")
PER=$(python -c "import json,sys; print(json.loads(open('$O/bench_plain_$R.log').read().strip().splitlines()[-1])['gpu_launches'])")
echo "skip $SKIP launches, capture $PER (one timed step)"
ncu --metrics gpu__time_duration.sum --clock-control none -s $SKIP -c $PER --csv --log-file $O/launches_$R.csv \
    python bench.py --steps 1 --warmup 3 --no-cpu-baseline > $O/bench_under_ncu_$R.log 2>&1
for t in gemm_fwd:svla_gemm_tc gemm_dgrad:svla_gemm_tc gemm_dgrad_mask:svla_gemm_tc gemm_wgrad:svla_gemm_tc attn:attn_tc_fwd attn:attn_tc_bwd \
         gae:gae_march loss:ppo_lag adam:clip_adam ln:layernorm_bwd; do
  tgt=${t%%:*}; k=${t##*:}
  ncu --set full --clock-control none --import-source on -k regex:$k -s 2 -c 1 -f -o $O/ncu_${R}_${tgt}_${k} \
      python tools/ncu_targets.py $tgt 4 > $O/ncu_${R}_${tgt}_${k}.log 2>&1
done
ls -la $O | tail -20
