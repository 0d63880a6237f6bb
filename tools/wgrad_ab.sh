# Round-1 A/B of the weight-gradient split-K policy.  The knobs it swept (CTAs per GEMM, minimum 64-row K blocks per split)
# are now the constants _WGRAD_TARGET / _WGRAD_MIN_KB / _WGRAD_BLOCK_N in kokoro_ruslan_b200/engine.py; to sweep again,
# edit them and run tools/gpu_ab.sh.  Results: engine._auto_splits docstring, DESIGN.md §7.
echo "see kokoro_ruslan_b200/engine.py:_auto_splits"
