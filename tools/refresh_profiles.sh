# Run on the GPU box (gpurun): the bench lines, ncu launch lists and the broad-phase ncu capture that profiles/ is built from.
set -x
cd $GRAFT_REPO_ROOT
python bench.py > gpurun_out/bench_default.json 2> gpurun_out/bench_default.err
python bench.py --workload cfg5 --steps 10 --warmup 3 > gpurun_out/bench_cfg5.json 2> gpurun_out/bench_cfg5.err
python bench.py --workload cfg2 --steps 50 --warmup 5 > gpurun_out/bench_cfg2.json 2> gpurun_out/bench_cfg2.err
python bench.py --impl reference > gpurun_out/bench_reference.json 2> gpurun_out/bench_reference.err
for c in cfg3 cfg5; do
  ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches_$c.csv python bench.py --workload $c --steps 2 --warmup 1 --no-cpu-baseline --no-e2e > gpurun_out/ncu_$c.log 2>&1
done
ncu --set full --clock-control none --import-source on -k 'regex:keys_kernel|tile_sums_kernel|scan_kernel|scatter_kernel' -c 12 -o gpurun_out/prof_broad_r01b -f python bench.py --workload cfg3 --steps 1 --warmup 1 --no-cpu-baseline --no-e2e > gpurun_out/ncu_broad.log 2>&1
tail -2 gpurun_out/ncu_broad.log
ls -la gpurun_out | tail -20
