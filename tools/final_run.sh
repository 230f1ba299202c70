# Regenerates the single-GPU evidence set under gpurun_out/ (copied into profiles/ afterwards): run on a B200 box from the repo root.
set -x
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu_final.txt 2>&1; tail -3 gpurun_out/pytest_gpu_final.txt
timeout 120 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke_final.txt 2>&1; tail -2 gpurun_out/smoke_final.txt
timeout 400 python bench.py --steps 50 --warmup 5 > gpurun_out/bench_c2_final.json 2> gpurun_out/bench_c2_final.err
timeout 200 python bench.py --config c1 --steps 200 --warmup 20 > gpurun_out/bench_c1_final.json 2> gpurun_out/bench_c1_final.err
timeout 200 python bench.py --config c3 --steps 200 --warmup 20 > gpurun_out/bench_c3_final.json 2> gpurun_out/bench_c3_final.err
timeout 300 python bench.py --config c4 --steps 20 --warmup 3 --quick > gpurun_out/bench_c4_final.json 2> gpurun_out/bench_c4_final.err
timeout 300 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/bench_reference_final.json 2> gpurun_out/bench_reference_final.err
python tools/show_bench.py gpurun_out/bench_c2_final.json gpurun_out/bench_c1_final.json gpurun_out/bench_c3_final.json gpurun_out/bench_c4_final.json gpurun_out/bench_reference_final.json
timeout 300 python tools/accuracy_report.py > gpurun_out/accuracy_final.txt 2>&1; tail -2 gpurun_out/accuracy_final.txt | cut -c1-300
AVD_NO_SIDE_STREAM=1 timeout 300 ncu --set full --clock-control none --import-source on -k regex:"fused3|wgrad3|dgrad3|unfold|learn_prep|adam|step_increment|actor_dm" --launch-skip 30 --launch-count 15 -o gpurun_out/r02_learn_step_final -f python tools/profile_learn.py 2 4096 1 > gpurun_out/ncu_learn_final.log 2>&1; tail -2 gpurun_out/ncu_learn_final.log
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r02_bench_launches.csv python bench.py --steps 2 --warmup 1 --quick --eager > gpurun_out/bench_under_ncu.log 2>&1
timeout 200 python tools/small_configs.py > gpurun_out/small_configs_final.txt 2>&1; cat gpurun_out/small_configs_final.txt
AVD_STAGE_TIMES=1 timeout 60 python tools/profile_learn.py 2 4096 1 2>&1 | tail -2 | head -1
timeout 60 python tools/profile_learn.py 2 4096 20
