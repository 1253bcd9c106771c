run() { # gpus threads percent
  B200LAP_HOST_NARROW_THREADS=$2 B200LAP_HOST_NARROW_PERCENT=$3 python -m torch.distributed.run --nnodes=1 --nproc-per-node $1 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus $1 --steps 16 --skip-cpu --skip-big --skip-configs > gpurun_out/r2_hybN$1_t$2_p$3.log 2> gpurun_out/r2_hybN$1_t$2_p$3.err
}
for cfg in $CFGS; do IFS=: read g t p <<< "$cfg"; run $g $t $p; done; echo done
