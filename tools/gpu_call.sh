mkdir -p gpurun_out
O=gpurun_out/r02b_e2e.txt; : > $O
for eager in 1 0; do for conn in "" 32; do for depth in 3 2; do
  if [ -n "$conn" ]; then export CUDA_DEVICE_MAX_CONNECTIONS=$conn; else unset CUDA_DEVICE_MAX_CONNECTIONS; fi
  DEPTH=$depth ORB_B200_EAGER_D2H=$eager timeout 120 python tools/e2e_probe.py >> $O 2>&1
done; done; done
export CUDA_DEVICE_MAX_CONNECTIONS=32
for lanes in 1 3; do ORB_B200_LANES=$lanes timeout 120 python tools/e2e_probe.py >> $O 2>&1; done
for chunk in 16 64; do ORB_B200_CHUNK=$chunk timeout 120 python tools/e2e_probe.py >> $O 2>&1; done
cat $O
