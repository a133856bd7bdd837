# Round-2 measurement batch (one gpurun call): driver-style bench, reference arm, 1024^2 config, ncu capture, launch list.
set -x
python bench.py --steps 20 --warmup 5 > gpurun_out/r02_bench_n1.json 2> gpurun_out/r02_bench_n1.err
python bench.py --impl reference --steps 5 --warmup 1 > gpurun_out/r02_bench_reference.json 2> gpurun_out/r02_bench_reference.err
python bench.py --size 1024 --per-gpu 128 --steps 5 --warmup 3 --est-images 0 --cpu-seconds 0 > gpurun_out/r02_bench_1024.json 2> gpurun_out/r02_bench_1024.err
python tools/one_pass.py 32 fp16x1_f8 2 && ncu --set full --clock-control none --import-source on -k regex:"conv_halo|upconv_res|first_conv" -s 12 -c 12 -o gpurun_out/r02_chain_final -f python tools/one_pass.py 32 fp16x1_f8 2 > gpurun_out/r02_ncu_chain_final.log 2>&1
python bench.py --steps 2 --warmup 3 --est-images 0 --cpu-seconds 0 > /dev/null 2>&1 && ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r02_launches_bench_step.csv python bench.py --steps 2 --warmup 3 --est-images 0 --cpu-seconds 0 > gpurun_out/r02_ncu_launches.log 2>&1
tail -c 600 gpurun_out/r02_bench_n1.err
