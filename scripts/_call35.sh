set -x
NXFX_LIB=networks_fenicsx_b200/csrc/libnxfx_b200_stamps.so timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29643 scripts/dist_stamps.py 21 2>&1 | grep -v "^W\|^\*\|OMP_NUM\|warn\|colors ="
