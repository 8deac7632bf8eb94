# The round-end evidence run of the final code (one GPU): tests, bench lines, launch lists and ncu captures (each ncu command after the same
# command ran plainly).  usage (on the GPU box): bash tools/capture_final.sh ; outputs under gpurun_out/ with the prefix r02f_
set -u
O=gpurun_out; P=r02f; mkdir -p $O
timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -8 > $O/${P}_gpu_tests.log
timeout 400 python bench.py > $O/${P}_bench_C3_1gpu.json 2> $O/${P}_bench_C3.err
for w in C1 C2 C4 C5; do timeout 200 python bench.py --workload $w --no-from-structs > $O/${P}_bench_${w}_1gpu.json 2> $O/${P}_bench_$w.err; done
timeout 300 python bench.py --impl reference > $O/${P}_bench_C3_reference_arm.json 2> $O/${P}_ref_C3.err
timeout 200 python bench.py --path group > $O/${P}_bench_group_1gpu.json 2> $O/${P}_group.err
timeout 200 python bench.py --path refine > $O/${P}_bench_refine_1gpu.json 2> $O/${P}_refine.err
timeout 200 python bench.py --path refine --impl reference > $O/${P}_bench_refine_reference_arm.json 2> $O/${P}_refine_ref.err
timeout 200 python __graft_entry__.py smoke > $O/${P}_smoke.log 2>&1
( cd tools/micro && for v in 1 3 4; do timeout 60 ./hamming_mma 200000 84 12 3 2 0 $v; done ) > $O/${P}_hamming_mma_prototype.jsonl 2>/dev/null
# launch lists (the commands above ran plainly first), then full captures of the step's main kernels
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file $O/${P}_launches_C3.csv python tools/prof_step.py C3 3 > $O/${P}_ncu1.log 2>&1
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file $O/${P}_launches_C4.csv python tools/prof_step.py C4 3 > $O/${P}_ncu2.log 2>&1
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 1200 --csv --log-file $O/${P}_launches_refine.csv python bench.py --path refine --steps 1 --warmup 1 --no-cpu-baseline > $O/${P}_ncu3.log 2>&1
timeout 300 ncu --set full --clock-control none --import-source on -k regex:"k_sweep_mma|k_pack_i8" -c 4 -o $O/prof_${P}_mma python tools/prof_step.py C4 2 > $O/${P}_ncu4.log 2>&1
timeout 300 ncu --set full --clock-control none --import-source on -k regex:"k_sweep$|k_pack_t2_2bit|k_phase_a|k_tier2|k_boruvka|k_entry_keys" -s 9 -c 9 -o $O/prof_${P}_c3 python tools/prof_step.py C3 2 > $O/${P}_ncu5.log 2>&1
timeout 300 ncu --set full --clock-control none --import-source on -k regex:"k_r_joins|k_r_doublets|k_r_link_exacts|k_r_columns|DeviceRadixSortOnesweep" -s 40 -c 8 -o $O/prof_${P}_refine python bench.py --path refine --steps 1 --warmup 0 --no-cpu-baseline > $O/${P}_ncu6.log 2>&1
echo done
