export ABEILLE_B200_KERNEL_TIMEOUT_S=60
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29517 bench.py --gpus 8 --steps 8 --warmup 4 > gpurun_out/t5b_bench_n8.json 2> gpurun_out/t5b_bench_n8.err
cat gpurun_out/t5b_bench_n8.json | python -c "import json,sys; d=json.loads(sys.stdin.read()); print('N=8 value %.4g ms/step %.2f e2e %.4g ranks_check %s'%(d['value'], d['ms_per_step'], d['e2e']['value'], d.get('ranks_check')))" || tail -20 gpurun_out/t5b_bench_n8.err
