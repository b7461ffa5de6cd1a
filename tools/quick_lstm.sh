# usage: bash tools/quick_lstm.sh "ENV=1 ENV2=2" ...   (one bench line per environment set)
for e in "$@"; do
  echo "== $e"
  env $e python bench.py --steps 10 --warmup 3 2>&1 | tail -1 | python -c "import json,sys; d=json.loads(sys.stdin.read()); print(d['ms_per_step'], 'lstm', d['stages_ms_per_step']['lstm_recurrence'])"
done
