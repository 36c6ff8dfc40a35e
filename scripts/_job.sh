set -x
export ABEILLE_B200_KERNEL_TIMEOUT_S=20
export ABEILLE_B200_EQ_STATS=1
timeout 500 python scripts/variant_probe.py > gpurun_out/t2d_variants.log 2>&1
grep -E "^==|gen3|event kernel" gpurun_out/t2d_variants.log | awk '/event kernel/{c++; if (c%4==0) print; next} {print}'
