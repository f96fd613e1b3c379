# A/B different builds of the library on the same box: each in its own process, interleaved twice
for rep in 1 2; do
for lib in "$@"; do
  echo "=== $lib (rep $rep)"
  if [ "$lib" = "HEAD" ]; then VARIANTS='[{}]' timeout 300 python tools_dev/ab_conv.py; else DS_LIB_PATH=$PWD/build/lib_$lib.so VARIANTS='[{}]' timeout 300 python tools_dev/ab_conv.py; fi
done
done
