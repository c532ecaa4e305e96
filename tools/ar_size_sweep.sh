TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29541"
for n in 2201600 550400 137600 8800000; do
for cfg in "0 16" "256 96"; do
  set -- $cfg
  N_FLOATS=$n NAFAE_AR_THREADS=$1 NAFAE_AR_CTAS=$2 P2P_BW=$([ $n = 137600 ] && echo 1) timeout 100 $TR tools/test_allreduce.py 2>&1 | grep -E "^world|rror|p2p" | sed "s/^/[n $n threads $1 ctas $2] /"
done
done
nvidia-smi topo -m | head -8
