cd $GRAFT_REPO_ROOT
for cl in 1 0; do for u in 1 64; do
echo "== cluster=$cl utts=$u"
B200_HMM_PROBE=1 B200_HMM_CLUSTER=$cl timeout 300 python bench_hmm.py --utts $u --frames 40 --warmup 8 --no-cpu-baseline 2>&1 | grep "hmm probe" | tail -2
done; done
