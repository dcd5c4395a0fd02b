for v in head g8 t16 t8s16 cur; do
  if [ $v = cur ]; then unset RTG_LIB; else export RTG_LIB=$PWD/build_variants/librt_$v.so; fi
  echo "== $v"
  timeout 300 python scripts/quick_perf.py synth256 "synth1024 4K a1" synth4096 2>&1 | grep -E "case|rror" | cut -c1-120
done
