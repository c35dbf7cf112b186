B="python bench.py --steps 40 --warmup 5 --no-extras --no-cpu-baseline"
pick() { python -c "import json,sys; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('$1', round(d['ms_per_step'],3), round(d['e2e']['value'],1), round(d['streaming']['ms_per_step'],3))"; }
timeout 600 python -m pytest tests -x -q -m gpu -k "replay or program or prefetch or golden" 2>&1 | tail -2
$B 2>/dev/null | pick new
SCN_OUT_MULTI=0 $B 2>/dev/null | pick no_multi
SCN_FIRST_PLAN_HERE=0 $B 2>/dev/null | pick no_firstplan
SCN_DECONV_EARLY=0 $B 2>/dev/null | pick no_deconv_early
SCN_OUT_MULTI=0 SCN_FIRST_PLAN_HERE=0 SCN_DECONV_EARLY=0 $B 2>/dev/null | pick old
$B 2>/dev/null | pick new
python tools/trace_forward.py bf16 > gpurun_out/s2_trace_digest.log 2>&1
