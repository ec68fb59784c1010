bash tools/gpu_ref_converged.sh c4 > gpurun_out/conv_c4.log 2>&1
cp gpurun_out/ref/c4_800x800_1000_ds5.npz tests/golden/ref_gpu/ 2>/dev/null
bash tools/gpu_round.sh tests,slots
