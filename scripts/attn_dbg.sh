for d in 0 15; do echo "== dbg $d"; DYT_ATTN_DBG=$d DYT_ATTN_TRACE=1 timeout 60 python scripts/attn_probe.py 2>&1 | grep -E "(wgA|wgB|mma) it=(10|11):"; done
