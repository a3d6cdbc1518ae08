for team in 0 1; do
  for sc in hi2q:1 hi2q:6 hi2q:12 bench:6; do
    echo -n "team=$team $sc: "
    UA_FUSED_TEAM=$team timeout 100 python tools/prof_one.py $sc --qubits 30 --reps 2 | tail -1
  done
  echo -n "team=$team c128 bench:6: "
  UA_FUSED_TEAM=$team timeout 100 python tools/prof_one.py bench:6 --qubits 29 --dtype c128 --reps 2 | tail -1
done
