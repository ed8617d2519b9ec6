timeout 120 python tools/attn_trace.py 2>&1 | tail -30
