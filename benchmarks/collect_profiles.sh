set -x
mkdir -p gpurun_out/r1
timeout 600 python -m pytest tests -q -m gpu --timeout 300 2>&1 | tail -3 > gpurun_out/r1/pytest_gpu.txt
timeout 300 python bench.py --steps 200 --warmup 10 > gpurun_out/r1/bench_default.json 2> gpurun_out/r1/bench_default.err
timeout 300 python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/r1/bench_reference.json 2> gpurun_out/r1/bench_reference.err
timeout 600 python bench.py --workload coco_c --steps 3 --warmup 1 > gpurun_out/r1/bench_coco_c.json 2> gpurun_out/r1/bench_coco_c.err
timeout 600 python bench.py --workload mpii_c --steps 3 --warmup 1 > gpurun_out/r1/bench_mpii_c.json 2> gpurun_out/r1/bench_mpii_c.err
timeout 300 python bench.py --workload advmix_mix --steps 50 --warmup 5 > gpurun_out/r1/bench_advmix_mix.json 2> gpurun_out/r1/bench_advmix_mix.err
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r1/launches.csv python bench.py --steps 3 --warmup 3 --no-cpu-baseline --no-graph > gpurun_out/r1/launches_bench.log 2>&1
timeout 300 ncu --set full --clock-control none --import-source on -k regex:"warp_affine_ws|heatmap_kernel" -s 8 -c 2 -o gpurun_out/r1/prof_crop_targets python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-graph > gpurun_out/r1/prof.log 2>&1
nvidia-smi --query-gpu=name,clocks.max.sm,clocks.max.mem,power.limit --format=csv > gpurun_out/r1/gpu.txt
ls -la gpurun_out/r1
timeout 300 python bench.py --workload jpeg_crop --steps 20 --warmup 3 > gpurun_out/r1/bench_jpeg_crop.json 2> gpurun_out/r1/bench_jpeg_crop.err
timeout 300 python tests/parity_report.py all > gpurun_out/r1/parity_report.txt 2>&1
timeout 300 python benchmarks/corruption_per_op.py > gpurun_out/r1/opbench.txt 2>&1

timeout 200 ncu --metrics gpu__time_duration.sum --clock-control none -k regex:jpeg_ -c 30 --csv --log-file gpurun_out/r1/launches_jpeg.csv python benchmarks/jpeg_decode_bench.py > gpurun_out/r1/launches_jpeg.log 2>&1
timeout 300 ncu --set full --clock-control none --import-source on -k regex:jpeg_huffman_parallel -s 2 -c 1 -o gpurun_out/r1/prof_jpeg_huffman python benchmarks/jpeg_decode_bench.py > gpurun_out/r1/prof_jpeg.log 2>&1
timeout 300 python benchmarks/jpeg_encode_bench.py > gpurun_out/r1/jpeg_encode_bench.txt 2>&1
timeout 400 ncu --set full --clock-control none --import-source on -k regex:"zoom_blur_smem|motion_blur_smem|gauss2d_rt|defocus_kernel|jpeg_entropy_encode|jpeg_fdct_quant" -c 10 -o gpurun_out/r1/prof_stencils -f python benchmarks/corruption_once.py zoom_blur,motion_blur,glass_blur,defocus_blur,jpeg_encode 3 > gpurun_out/r1/prof_stencils.log 2>&1
