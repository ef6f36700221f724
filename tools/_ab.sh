for q in 0 1; do
  XDB200_GN_QSTATS=$q python bench.py --workload unet --steps 3 --warmup 3 --no-cpu > gpurun_out/y_unet_qs$q.log 2>&1
  echo "qstats $q: $(grep -o '"value": [0-9.]*' gpurun_out/y_unet_qs$q.log | head -1) $(grep -o '"ms_per_step": [0-9.]*' gpurun_out/y_unet_qs$q.log | head -1) $(grep -o '"gpu_launches": [0-9]*' gpurun_out/y_unet_qs$q.log | head -1)"
done
