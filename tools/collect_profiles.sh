#!/bin/bash
# Profile evidence for one round (run under gpurun, 1 GPU):  bash tools/collect_profiles.sh r01
# 1. ncu launch list over ~one step of the exact bench command   2. ncu --set full of the roofline kernels
R=${1:-r02}
O=gpurun_out
mkdir -p $O
python bench.py --steps 1 --warmup 3 --no-cpu-baseline --no-library-baseline > $O/bench_plain_$R.log 2>&1
SKIP=$(grep -o "launches before the timed region: [0-9]*" $O/bench_plain_$R.log | grep -o "[0-9]*$")
PER=$(python -c "import json,sys; print(json.loads(open('$O/bench_plain_$R.log').read().strip().splitlines()[-1])['gpu_launches'])")
echo "skip $SKIP launches, capture $PER (one timed step)"
ncu --metrics gpu__time_duration.sum --clock-control none -s $SKIP -c $PER --csv --log-file $O/launches_$R.csv \
    python bench.py --steps 1 --warmup 3 --no-cpu-baseline --no-library-baseline > $O/bench_under_ncu_$R.log 2>&1
if [ -n "$LAUNCHES_ONLY" ]; then  # refresh the launch list only (the ncu --set full captures take ~8 GPU-minutes)
  python tools/ncu_summarize.py launches $O/launches_$R.csv $O/launches_$R.json > $O/launches_$R.txt 2>&1
  gzip -f $O/launches_$R.csv
  exit 0
fi
for t in gemm_fwd:svla_gemm_tc gemm_fwd_bits:svla_gemm_tc gemm_dgrad:svla_gemm_tc gemm_dgrad_bits:svla_gemm_tc \
         gemm_res:svla_gemm_tc gemm_wgrad:svla_gemm_tc gemm_x3:svla_gemm_tc attn:attn_ws_fwd attn:attn_ws_bwd \
         attn_x3:attn_ws_fwd attn_x3:attn_ws_bwd_x3 attn_drop:attn_ws_bwd split:split_concat \
         gae:gae_march loss:ppo_lag adam:clip_adam ln:layernorm_bwd ln_fwd:layernorm_fwd attn_cls:attn_cls_fwd \
         attn_cls:attn_cls_bwd vision:attn_flash_fwd; do
  tgt=${t%%:*}; k=${t##*:}
  ncu --set full --clock-control none --import-source on -k regex:$k -s 2 -c 1 -f -o $O/ncu_${R}_${tgt}_${k} \
      python tools/ncu_targets.py $tgt 4 > $O/ncu_${R}_${tgt}_${k}.log 2>&1
done
# summaries are what travels back (gpurun merges at most 64 MiB): full reports only of the kernels read at source level
python tools/ncu_summarize.py launches $O/launches_$R.csv $O/launches_$R.json > $O/launches_$R.txt 2>&1
python tools/ncu_summarize.py rep $O/ncu_${R}_*.ncu-rep > $O/ncu_$R.md 2>/dev/null
python tools/ncu_summarize.py traffic $O/ncu_${R}_traffic.json $O/ncu_${R}_*.ncu-rep > /dev/null 2>&1
gzip -f $O/launches_$R.csv
for f in $O/ncu_${R}_*.ncu-rep; do
  case "$f" in *attn_attn_ws_fwd*|*attn_attn_ws_bwd*|*gemm_res*) ;; *) rm -f "$f" ;; esac
done
du -sh $O; ls -la $O | tail -30
