for t in grab1 "" grab16; do
  if [ -z "$t" ]; then L=""; else L="PN_LIB=$PWD/pienerf_b200/lib/libpienerf_b200_$t.so"; fi
  echo "== variant ${t:-default}"; env $L timeout 200 python scripts/mode_compare.py 3 1.0 2>&1 | tail -1 | cut -c1-110; env $L timeout 200 python scripts/mode_compare.py 3 1.0 trex 2>&1 | tail -1 | cut -c1-110
done
