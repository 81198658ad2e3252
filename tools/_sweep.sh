python -m pytest tests/test_gpu_kernels.py -m gpu -q -k "elementwise or ragged or sync" 2>&1 | tail -2
python tools/kernel_profile.py 60 --no-overlap 2>&1 | grep "bn_bwd_apply\|stem_im2col\|sum of kernel"
CSTP_BN_BWD_OCC=4 python tools/kernel_profile.py 60 --no-overlap 2>&1 | grep "bn_bwd_apply\|stem_im2col\|sum of kernel"
python tools/step_time.py 60 | tail -1
CSTP_BN_BWD_OCC=4 python tools/step_time.py 60 | tail -1
