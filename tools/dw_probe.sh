cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
timeout 300 python tools/halo_sw_probe.py stream 2>&1 | grep -E "^TIME|^STREAM|Error|error|Traceback" | head -60
