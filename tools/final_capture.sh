set -x
timeout 400 python -m pytest tests -x -q -m gpu 2>&1 | tail -6 > gpurun_out/h1_tests.txt
timeout 120 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/h1_smoke.txt 2>&1; echo "smoke rc $?" >> gpurun_out/h1_smoke.txt
timeout 300 python bench.py > gpurun_out/h1_bench_n1.json 2> gpurun_out/h1_bench.err
timeout 200 python bench.py --config cfg3 --steps 200 --no-cpu-baseline --no-serving > gpurun_out/h1_bench_cfg3.json 2> gpurun_out/h1_bench_cfg3.err
timeout 100 python tools/step_timeline.py > gpurun_out/h1_timeline_n1.txt 2>&1
timeout 200 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/h1_launches.csv python bench.py --steps 2 --warmup 1 --no-graph --no-cpu-baseline --no-serving --no-extras > gpurun_out/h1_ncu_launch.log 2>&1
timeout 300 ncu --set full --import-source on --clock-control none -k regex:"optimizer_step|tower_mlp2|retrieval_fwd_dq|retrieval_bwd|dq_finalize|sparse_prepare" --launch-skip 21 -c 7 -o gpurun_out/h1_step python bench.py --steps 2 --warmup 3 --no-graph --no-cpu-baseline --no-serving --no-extras > gpurun_out/h1_ncu_step.log 2>&1
cat gpurun_out/h1_tests.txt; tail -3 gpurun_out/h1_smoke.txt; ls -la gpurun_out/h1_*
