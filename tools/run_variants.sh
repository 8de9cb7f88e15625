run() { python bench.py --no-cpu --no-e2e "$@" 2>/dev/null | tail -1 | python -c "import sys,json; d=json.loads(sys.stdin.read()); print(round(d['value']), round(d['ms_per_step'],4), round(d['step_latency_ms'],4), round(d['roofline']['kernel_ms_per_launch'],4))"; }
echo "p1"; run --pipeline 1
echo "p2 share"; run --pipeline 2
echo "p2 noshare"; run --pipeline 2 --no-share
echo "p3 noshare"; run --pipeline 3 --no-share
echo "p2 share token"; OOKD_SCREEN_TOKEN=1 run --pipeline 2
echo "p2 share steps30"; run --pipeline 2 --steps 30
