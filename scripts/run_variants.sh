for v in "$@"; do
  if [ $v = cur ]; then unset RTG_LIB; else export RTG_LIB=$PWD/build_variants/librt_$v.so; fi
  echo "== $v"
  timeout 300 python scripts/quick_perf.py ${CASES:-synth} 2>&1 | grep -E "case|rror" | cut -c1-125
done
