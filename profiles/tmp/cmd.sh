timeout -s KILL 900 python -m pytest tests -q -m gpu -x 2>&1 | tail -3
timeout -s KILL 240 python profiles/run_matrix.py 22 > gpurun_out/matrix_final.jsonl 2> gpurun_out/matrix_final.err
timeout -s KILL 120 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -1
timeout -s KILL 600 python profiles/fuzz_parity.py 1500 78 > gpurun_out/fuzz_single.jsonl 2> gpurun_out/fuzz_single.err; tail -n 1 gpurun_out/fuzz_single.jsonl
timeout -s KILL 300 ncu --set full --clock-control none --import-source on -k regex:smolyak_leaf -c 1 -s 1 -o gpurun_out/prof_leaf_d3_single -f python profiles/run_highdim_one.py 3 3 4194304 > gpurun_out/ncu_d3.log 2>&1; tail -n 2 gpurun_out/ncu_d3.log
timeout -s KILL 400 python bench.py > gpurun_out/bench_final.json 2> gpurun_out/bench_final.err; cat gpurun_out/bench_final.json | cut -c1-300
