set -x
export ABEILLE_B200_KERNEL_TIMEOUT_S=20
export ABEILLE_B200_EQ_STATS=1
timeout 700 python scripts/variant_probe.py > gpurun_out/t1d_variants.log 2>&1
grep -E "^==|gen3|event kernel" gpurun_out/t1d_variants.log | awk '/event kernel/{c++; if (c%4==0) print; next} {print}'
unset ABEILLE_B200_EQ_STATS
timeout 600 python -m pytest tests -m gpu -x -q > gpurun_out/t1d_pytest.log 2>&1; echo "pytest exit $?" >> gpurun_out/t1d_pytest.log
tail -5 gpurun_out/t1d_pytest.log
