D=gpu_ray_tracing_for_waveguide_based_ar_display_b200
cp $D/libwgrt.so /tmp/libwgrt_base.so
for v in W P base; do
  if [ $v = base ]; then cp /tmp/libwgrt_base.so $D/libwgrt.so; else cp $D/csrc/build/libwgrt_$v.so $D/libwgrt.so; fi
  echo "variant $v"; python tools/quick_perf.py --rays 5000 --iters 4 2>&1 | tail -1 | cut -c1-110
done
