export ABEILLE_B200_KERNEL_TIMEOUT_S=120
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q -k "exact or branchless" 2>&1 | tail -5
timeout 500 python scripts/modes_probe.py > gpurun_out/t9i_modes_probe.json 2> gpurun_out/t9i_modes_probe.err; echo rc=$?; cat gpurun_out/t9i_modes_probe.json | cut -c1-1500; tail -3 gpurun_out/t9i_modes_probe.err
