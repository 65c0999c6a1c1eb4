for e in ${EXPS:-0 2 64 128 8 16}; do echo "== OSB_TA_EXP=$e"; OSB_TA_EXP=$e ONLY="attn tiles" timeout 100 python tests/attn_tiles_prof.py 10 2>&1 | tail -3; done
