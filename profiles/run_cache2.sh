timeout 1700 python -m pytest tests -m gpu -x -q 2>&1 | tail -2
python bench.py --steps 1000 --warmup 10 --no-cpu-baseline --no-plain 2>/dev/null | python -c "import sys,json; d=json.loads(sys.stdin.read()); a=d['assembly']; print(d['value'], d['e2e']['value'], {k:a[k] for k in a if k!='dict_api'})"
