# One GPU call for the auto-encoder kernels: decoder parity tests, encode benchmark, ncu capture.  usage: bash tools/decoder_round.sh <tag>
mkdir -p gpurun_out
tag=${1:-e1}
timeout 600 python -m pytest tests/test_gpu_decoder.py -m gpu -x -q > gpurun_out/pytest_dec_$tag.log 2>&1; tail -3 gpurun_out/pytest_dec_$tag.log
timeout 300 python tools/bench_decoder.py --encode > gpurun_out/enc_$tag.log 2> gpurun_out/enc_$tag.err; tail -c 300 gpurun_out/enc_$tag.err; cat gpurun_out/enc_$tag.log
timeout 600 ncu --set full --clock-control none --import-source on -k regex:'k_enc_' -c 2 -o gpurun_out/enc_$tag -f python tools/bench_decoder.py --encode --iters 1 --no-cpu-baseline > gpurun_out/ncu_enc_$tag.log 2>&1; tail -2 gpurun_out/ncu_enc_$tag.log
