set -x
export ABEILLE_B200_KERNEL_TIMEOUT_S=20
timeout 60 python scripts/profile_target.py c5g7_delta_collision_fullmesh.yaml 10000000 4 2>&1 | grep -E "gen3"
timeout 60 python scripts/profile_target.py c5g7_delta_collision_fullmesh.yaml 10000000 4 2>&1 | grep -E "gen3"
