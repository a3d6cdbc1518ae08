for th in 128 768; do
  echo -n "threads=$th hi2q:9 "
  UA_FUSED_THREADS=$th timeout 200 python tools/prof_one.py hi2q:9 --qubits 30 --reps 2 | tail -1
  echo -n "threads=$th hi2q:1 "
  UA_FUSED_THREADS=$th timeout 200 python tools/prof_one.py hi2q:1 --qubits 30 --reps 2 | tail -1
  echo -n "threads=$th bench:6 "
  UA_FUSED_THREADS=$th timeout 200 python tools/prof_one.py bench:6 --qubits 30 --reps 2 | tail -1
  echo -n "threads=$th c128 bench:6 "
  UA_FUSED_THREADS=$th timeout 200 python tools/prof_one.py bench:6 --qubits 29 --dtype c128 --reps 2 | tail -1
done
