# same-box A/B: dual pipelines also for N = 256 layers (MG_DUAL=2) vs thin-N only (default)
for d in 1 2; do echo "== MG_DUAL=$d"; MG_DUAL=$d python tools/whatif_spade.py 2>&1 | grep -E "(f16|tf32) +MG_DBG= 0"; for cfg in "bf3 256 128 256" "tf32 128 256 512" "f16 128 256 512" "bf3 512 512 64"; do MG_DUAL=$d python tools/prof_conv.py $cfg | grep "halo=0 MG_DBG=0"; done; done
