for a in "" "--trained" "--flush" "--trained --flush" "--user-begin 3000000"; do
  echo "== $a"; python scripts/prof_pass1.py $a 2>/dev/null | tail -1 | python -c "import json,sys; d=json.loads(sys.stdin.read()); print(d['ms'], d['user_steps_per_s'], d['organic_draws'], d['phase_share'])"
done
