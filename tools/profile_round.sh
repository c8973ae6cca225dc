set -x
NCU="ncu --metrics gpu__time_duration.sum --clock-control none --csv"
timeout 900 $NCU --log-file gpurun_out/r02_bench_launches.csv python bench.py --profile-run --steps 2 --warmup 1 --batches 2 --no-cpu-baseline --no-configs --no-device-extract > gpurun_out/r02_bench_profile_run.json 2> gpurun_out/r02_bench_profile_run.err
timeout 300 $NCU --log-file gpurun_out/r02_pruned_step_launches.csv python tools/step_profile.py --steps 3 --warmup 2 > gpurun_out/r02_pruned_step.txt 2>&1
timeout 300 $NCU --log-file gpurun_out/r02_full_formulation_launches.csv python tools/step_profile.py --full --steps 2 --warmup 1 > gpurun_out/r02_full_step.txt 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:"pack_w_pair|hub_prepass|gcn_layer_fwd_pair" -s 9 -c 3 -o gpurun_out/r02_layer_full -f python bench.py --roofline-only > gpurun_out/r02_layer_full.txt 2>&1
ls -la gpurun_out/ | tail -12
