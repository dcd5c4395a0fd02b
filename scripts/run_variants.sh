for v in "$@"; do
  lib=${v%%:*}; opts=""; [ "$lib" != "$v" ] && opts=${v#*:}
  if [ $lib = cur ]; then unset RTG_LIB; else export RTG_LIB=$PWD/build_variants/librt_$lib.so; fi
  echo "== $v"
  RTG_OPTS=$opts timeout 300 python scripts/quick_perf.py ${CASES:-synth} 2>&1 | grep -E "case|rror" | cut -c1-125
done
