# Run on the GPU box (gpurun): the bench lines, ncu launch lists and ncu captures that profiles/ is built from.
#   gpurun --timeout 2400 -- 'bash tools/refresh_profiles.sh'    then, here:    python tools/update_profiles.py r02
set -x
cd $GRAFT_REPO_ROOT
O=gpurun_out
python bench.py > $O/bench_default.json 2> $O/bench_default.err
python bench.py --impl reference > $O/bench_reference.json 2> $O/bench_reference.err
python tools/small_world_latency.py > $O/small_world_latency.txt 2>&1
python tools/hub_scene.py > $O/hub_scene.txt 2>&1
python tools/hub_scene.py 100000 400000 static >> $O/hub_scene.txt 2>&1
python tools/dense_longrun.py 1048576 100 > $O/dense_longrun.txt 2>&1
for c in cfg3 cfg5; do
  python tools/collide_bench.py $c > $O/collide_bench_$c.txt 2>&1
  ncu --metrics gpu__time_duration.sum --clock-control none -c 300 --csv --log-file $O/launches_collide_$c.csv python tools/collide_bench.py $c 1048576 4 > $O/ncu_l_$c.log 2>&1
done
ncu --metrics gpu__time_duration.sum --clock-control none -c 300 --csv --log-file $O/launches_bench_cfg3.csv python bench.py --workload cfg3 --steps 2 --warmup 1 --no-cpu-baseline --no-e2e > $O/ncu_l_bench.log 2>&1
# full captures, one launch each (launch 9 of the kernel = a timed step after warm-up)
ncu --set full --clock-control none --import-source on -k regex:narrow_kernel -s 8 -c 1 -o $O/prof_narrow_cfg3 -f python tools/collide_bench.py cfg3 1048576 6 > $O/ncu_f1.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:narrow_kernel -s 8 -c 1 -o $O/prof_narrow_cfg5 -f python tools/collide_bench.py cfg5 1048576 6 > $O/ncu_f2.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:prep_kernel -s 8 -c 1 -o $O/prof_prep_cfg5 -f python tools/collide_bench.py cfg5 1048576 6 > $O/ncu_f3.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:sparse_step_kernel -s 8 -c 1 -o $O/prof_sparse_cfg3 -f python tools/collide_bench.py cfg3 1048576 6 > $O/ncu_f4.log 2>&1
TR_NO_SPARSE=1 ncu --set full --clock-control none --import-source on -k regex:prep_kernel -s 8 -c 1 -o $O/prof_prep_cfg3 -f python tools/collide_bench.py cfg3 1048576 6 > $O/ncu_f4b.log 2>&1
ncu --set full --clock-control none --import-source on -k 'regex:keys_kernel|tile_sums_kernel|scan_kernel|scatter_kernel' -s 32 -c 4 -o $O/prof_broad_cfg3 -f python tools/collide_bench.py cfg3 1048576 6 > $O/ncu_f5.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:response_flow_kernel -s 8 -c 1 -o $O/prof_flow_cfg5 -f python tools/collide_bench.py cfg5 1048576 6 > $O/ncu_f6.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:gravity_kernel -s 1 -c 1 -o $O/prof_gravity_cfg3 -f python bench.py --workload cfg3 --steps 1 --warmup 1 --no-cpu-baseline --no-e2e --no-secondary > $O/ncu_f7.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:tiny_step_kernel -s 2 -c 1 -o $O/prof_tiny -f python tools/small_world_latency.py > $O/ncu_f8.log 2>&1
ls -la $O | tail -30
