for v in "" "MGB_NO_PDL=1" "MGB_DIST_NOROT=1" "MGB_NO_PDL=1 MGB_DIST_NOROT=1"; do
  echo "== variant: $v"
  env $v timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29512 scratch/peer_timeline.py 2>&1 | grep -E "begin_us" | tr -d ' \n'; echo
done
