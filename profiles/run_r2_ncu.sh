#!/bin/bash
# Round 2: ncu captures whose summaries are kept under profiles/ (run under gpurun, ONE GPU).
# Each capture follows a plain run of the same program that exited 0.
set -x
O=gpurun_out
NCU="ncu --set full --clock-control none --import-source on -f"
python profiles/run_kernel.py 4194304 2 > $O/r2_run_kernel.log 2>&1 && $NCU -k regex:smolyak_leaf_kernel -c 1 -o $O/r2_leaf python profiles/run_kernel.py 4194304 1 > $O/r2_ncu_leaf.log 2>&1
python profiles/run_basis_one.py grad 100000 > /dev/null 2>&1 && $NCU -k regex:smolyak_tile_basis -s 1 -c 1 -o $O/r2_basis_grad python profiles/run_basis_one.py grad 100000 > $O/r2_ncu_bg.log 2>&1
python profiles/run_basis_one.py values 400000 > /dev/null 2>&1 && $NCU -k regex:smolyak_tile_basis -s 1 -c 1 -o $O/r2_basis_values python profiles/run_basis_one.py values 400000 > $O/r2_ncu_bv.log 2>&1
python profiles/run_uni.py 67108864 2 2 > $O/r2_run_uni.log 2>&1 && $NCU -k regex:uni_lincomb_kernel -s 1 -c 1 -o $O/r2_uni python profiles/run_uni.py 67108864 2 2 > $O/r2_ncu_uni.log 2>&1
python profiles/run_general.py 524288 > $O/r2_run_general.log 2>&1 && $NCU -k regex:smolyak_general_kernel -s 1 -c 1 -o $O/r2_general python profiles/run_general.py 524288 > $O/r2_ncu_general.log 2>&1
# launch list of the bench command (per-launch durations; cold-cache, serialised)
python bench.py --steps 2 --warmup 3 --no-cpu-baseline > $O/r2_bench_plain.json 2> $O/r2_bench_plain.err && ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file $O/r2_bench_launches.csv python bench.py --steps 2 --warmup 3 --no-cpu-baseline > $O/r2_bench_under_ncu.log 2>&1
ls -la $O/*.ncu-rep $O/r2_bench_launches.csv
