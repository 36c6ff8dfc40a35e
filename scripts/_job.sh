set -x
export ABEILLE_B200_KERNEL_TIMEOUT_S=20
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/s1g_pytest.log 2>&1; echo "pytest exit $?" >> gpurun_out/s1g_pytest.log
tail -8 gpurun_out/s1g_pytest.log
timeout 600 python scripts/variant_probe.py > gpurun_out/s1g_variants.log 2>&1
grep -E "^==|gen3" gpurun_out/s1g_variants.log
ABEILLE_B200_NO_SMEM_TABLES=1 timeout 600 python scripts/variant_probe.py > gpurun_out/s1g_variants_notables.log 2>&1
grep -E "^==|gen3" gpurun_out/s1g_variants_notables.log
