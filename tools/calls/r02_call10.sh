mkdir -p gpurun_out
for i in 1 2; do
for pol in all auto none; do echo "policy $pol"; CSTP_FUSE_POLICY=$pol timeout 300 python tools/step_time.py 60 --steps 20 2>/dev/null | tail -1 | cut -c1-90; done
done
