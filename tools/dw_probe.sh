cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
for m in 0 1 5; do Y11_HALO_SW=$m timeout 200 python tools/halo_sw_probe.py time 2>&1 | grep -E "^TIME|Error|error" | head -20; done
