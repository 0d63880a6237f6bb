for n in 1 2 4; do echo "KR_NORM_BLOCKS_PER_SM=$n"; KR_NORM_BLOCKS_PER_SM=$n python tools/norm_bench.py | grep -E "layernorm_bwd|rmsnorm"; done
for n in 2 3 6 12; do echo "KR_PREP_BWD_BLOCKS_PER_SM=$n"; KR_PREP_BWD_BLOCKS_PER_SM=$n python tools/norm_bench.py | grep "qkv_prep_bwd"; done
for n in 2 4 8 16; do echo "KR_PREP_FWD_BLOCKS_PER_SM=$n"; KR_PREP_FWD_BLOCKS_PER_SM=$n python tools/norm_bench.py | grep "qkv_prep_fwd"; done
