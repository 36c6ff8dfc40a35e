export ABEILLE_B200_KERNEL_TIMEOUT_S=60
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29517 bench.py --gpus 8 --steps 8 --warmup 4 --no-ranks-check > gpurun_out/t5c_bench_n8.json 2> gpurun_out/t5c_bench_n8.err
python -c "import json; d=json.load(open('gpurun_out/t5c_bench_n8.json')); print('N=8 value %.4g ms/step %.2f e2e %.4g placement %s'%(d['value'], d['ms_per_step'], d['e2e']['value'], d.get('host_placement')))" || tail -20 gpurun_out/t5c_bench_n8.err
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29518 bench.py --gpus 8 --config 4 --steps 4 --warmup 3 --no-parity > gpurun_out/t5c_bench_config4_n8.json 2> gpurun_out/t5c_bench_config4_n8.err
python -c "import json; d=json.load(open('gpurun_out/t5c_bench_config4_n8.json')); print('config4 N=8 value %.4g ms/step %.2f kernel %.2f n/gen %d'%(d['value'], d['ms_per_step'], d['roofline']['kernel_ms'], d['config']['particles_per_generation']))" || tail -20 gpurun_out/t5c_bench_config4_n8.err
lscpu | grep -E "NUMA|Socket|Model name|^CPU\(s\)" | head -8
nvidia-smi topo -m | head -14
