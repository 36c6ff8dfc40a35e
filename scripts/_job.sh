export ABEILLE_B200_KERNEL_TIMEOUT_S=120
timeout 500 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29519 bench.py --gpus 8 --config 5 --steps 1 --warmup 0 --no-parity > gpurun_out/t6b_bench_config5_n8.json 2> gpurun_out/t6b_bench_config5_n8.err
python -c "import json; d=json.load(open('gpurun_out/t6b_bench_config5_n8.json')); print('config5 N=8 value %.4g ms/step %.1f'%(d['value'], d['ms_per_step']), d['config']['noise_generations_per_batch'], d['config']['noise_particles_per_batch'])" || tail -8 gpurun_out/t6b_bench_config5_n8.err
