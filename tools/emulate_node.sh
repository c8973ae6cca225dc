#!/bin/bash
# One rank of an 8-rank job on a 1-GPU box: the bench on 4 cores, seven host-side stand-ins (tools/host_contention.py)
# on 4 cores each beside it.  Usage: bash tools/emulate_node.sh <tag> <device_finish 0|1>
tag=${1:-emu}; df=${2:-0}
n=$(nproc); per=4
slim=""; [ "$df" = "1" ] && slim="--slim"
pids=""
for r in 1 2 3 4 5 6 7; do
  lo=$((r * per)); hi=$((lo + per - 1))
  [ $hi -ge $n ] && break
  taskset -c $lo-$hi timeout 120 python tools/host_contention.py --seconds 70 $slim > gpurun_out/${tag}_standin$r.log 2>&1 &
  pids="$pids $!"
done
sleep 12
LOCAL_WORLD_SIZE=8 GMETA_B200_DEVICE_FINISH=$df taskset -c 0-$((per - 1)) timeout 200 python bench.py --no-cpu-baseline --no-configs \
  --no-device-extract > gpurun_out/${tag}.json 2> gpurun_out/${tag}.err
for p in $pids; do wait $p; done
cat gpurun_out/${tag}_standin*.log | grep host_contention
