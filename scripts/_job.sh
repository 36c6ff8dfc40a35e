set -x
export ABEILLE_B200_KERNEL_TIMEOUT_S=20
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/t1a_pytest.log 2>&1; echo "pytest exit $?" >> gpurun_out/t1a_pytest.log
tail -15 gpurun_out/t1a_pytest.log
timeout 300 python -c "
import sys; sys.path.insert(0,'.')
from scripts.perf_probe import run
run('c5g7_delta_collision_fullmesh.yaml', 10000000, 4, True)
" > gpurun_out/t1a_perf.log 2>&1
grep gen gpurun_out/t1a_perf.log
ABEILLE_B200_STAGED=1 timeout 300 python -c "
import sys; sys.path.insert(0,'.')
from scripts.perf_probe import run
run('c5g7_delta_collision_fullmesh.yaml', 10000000, 4, True)
" > gpurun_out/t1a_perf_staged.log 2>&1
grep gen gpurun_out/t1a_perf_staged.log
