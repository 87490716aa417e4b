for amp in 0 1 2 4 8 16 32; do
  timeout 90 python tools/prof_level.py $amp 2 > /tmp/lp_$amp.txt 2>&1; echo "amp $amp rc=$? $(tail -1 /tmp/lp_$amp.txt)"
done
