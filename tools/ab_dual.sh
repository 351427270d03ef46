# same-box A/B of the dual (two producer/issuer pairs) mode on thin-N layers
for d in 0 1; do echo "== MG_DUAL=$d"; for cfg in "bf3 128 64 512" "bf3 64 64 512" "bf3 256 128 256" "tf32 256 128 256" "tf32 128 64 512" "f16 128 128 512"; do MG_DUAL=$d python tools/prof_conv.py $cfg | grep "halo=0 MG_DBG=0"; done; done
echo "== halo + dual"; MG_HALO=1 python tools/prof_conv.py bf3 128 64 512 | grep "halo=1 MG_DBG=0"
