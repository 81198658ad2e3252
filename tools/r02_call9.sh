mkdir -p gpurun_out
for c in c2t c2s; do
for d in 0 4; do echo "dbg=$d"; CSTP_WG_DBG=$d CSTP_HC_DBG=$d python tools/prologue_probe.py $c 1 60 2>&1 | grep -v dgrad; done
python tools/prologue_probe.py $c 0 60 2>&1 | grep -v dgrad
done
