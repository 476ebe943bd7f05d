for i in 1 2; do
for v in 0 6 8; do echo -n "dbg=$v bits: "; SVLA_TC_DBG=$v python tools/ncu_targets.py gemm_fwd_bits 24 | tail -1 | python -c "
import sys,ast; l=sys.stdin.read(); xs=ast.literal_eval(l[l.index('['):]); xs=xs[4:]; print(round(sum(xs)/len(xs)*1000,1),'us')"; done
echo -n "plain relu: "; python tools/ncu_targets.py gemm_fwd 24 | tail -1 | python -c "
import sys,ast; l=sys.stdin.read(); xs=ast.literal_eval(l[l.index('['):]); xs=xs[4:]; print(round(sum(xs)/len(xs)*1000,1),'us')"
done
