set -x
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv
for ls in unset 0 1; do
  for nb in 1 2; do
    if [ $ls = unset ]; then unset SPIRK_V3_LOCKSTEP; else export SPIRK_V3_LOCKSTEP=$ls; fi
    echo "== LOCKSTEP=$ls nb=$nb"
    python tools/bench_vmult.py --refine 6 --variants 0 --reps 20 --nb $nb
  done
done
unset SPIRK_V3_LOCKSTEP
echo "== r7 default"; python tools/bench_vmult.py --refine 7 --variants 0 --reps 10 --nb 1
export SPIRK_V3_LOCKSTEP=0; echo "== r7 lockstep0"; python tools/bench_vmult.py --refine 7 --variants 0 --reps 10 --nb 1
unset SPIRK_V3_LOCKSTEP
echo "== r5"; python tools/bench_vmult.py --refine 5 --variants 0 --reps 20 --nb 2
