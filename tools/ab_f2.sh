for cfg in "128 12" "256 12" "128 11" "512 12"; do
  set -- $cfg
  echo "c128 threads=$1 tile=$2"
  UA_FUSED_THREADS=$1 UA_TILE_BITS=$2 timeout 200 python tools/prof_one.py bench:6 --qubits 29 --dtype c128 --reps 2 | tail -2 | tr '\n' ' '; echo
done
