# The round-end evidence run: tests, bench lines, shard probes, launch lists and ncu captures (each ncu command after the same command ran plainly).
# usage (on the GPU box): bash tools/capture_round.sh ; outputs under gpurun_out/
set -u
O=gpurun_out; mkdir -p $O
timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -25 > $O/r02z_tests.log
timeout 400 python bench.py > $O/r02z_bench_C3.json 2> $O/r02z_bench_C3.err
for w in C1 C2 C4 C5; do timeout 300 python bench.py --workload $w --no-from-structs > $O/r02z_bench_$w.json 2> $O/r02z_bench_$w.err; done
timeout 300 python bench.py --impl reference > $O/r02z_ref_C3.json 2> $O/r02z_ref_C3.err
timeout 200 python bench.py --path group > $O/r02z_group.json 2> $O/r02z_group.err
timeout 200 python bench.py --path group --impl reference > $O/r02z_group_ref.json 2> $O/r02z_group_ref.err
for c in 4 8 16; do echo "EJOIN_SHARD_CHUNKS=$c"; EJOIN_SHARD_CHUNKS=$c timeout 200 python tools/shard_probe.py 8 C3; done > $O/r02z_probe8.log 2>&1
timeout 200 python tools/shard_probe.py 2 C3 > $O/r02z_probe2.log 2>&1
timeout 200 python __graft_entry__.py smoke > $O/r02z_smoke.log 2>&1
# launch list of the bench command (after it ran plainly above), then full captures of the step's main kernels
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 1500 --csv --log-file $O/r02z_launches_bench.csv python bench.py --steps 2 --warmup 1 --no-cpu-baseline --no-from-structs > $O/r02z_ncu0.log 2>&1
timeout 200 python tools/prof_step.py C3 3 > $O/r02z_plain.log 2>&1 && timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file $O/r02z_launches.csv python tools/prof_step.py C3 3 > $O/r02z_ncu1.log 2>&1
timeout 400 ncu --set full --clock-control none --import-source on -k regex:"k_sweep|k_pack_t2_2bit|k_phase_a|k_phaseb_split|k_tier2|k_boruvka" -s 14 -c 14 -o $O/prof_r02z python tools/prof_step.py C3 3 > $O/r02z_ncu2.log 2>&1
timeout 200 ncu --set full --clock-control none --import-source on -k regex:k_group_pairs -s 1 -c 1 -o $O/prof_r02z_group python tools/group_bench.py --n 20000 --cpu-sample 0 --reps 2 > $O/r02z_ncu3.log 2>&1
echo done
