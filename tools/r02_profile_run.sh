set -x
mkdir -p gpurun_out/r02
cd tools/tc_probe
{ echo "== onehot 33"; ./tc_probe onehot 33; echo "== random 33 (big, several items per CTA)"; SP_PROBE_BIG=1 ./tc_probe random 33; echo "== random 31 (bf16 mode)"; SP_PROBE_BIG=1 ./tc_probe random 31; echo "== random 32 (generation 2)"; SP_PROBE_BIG=1 ./tc_probe random 32; echo "== stress 33"; SP_PROBE_BIG=1 SP_PROBE_REPEAT=300 ./tc_probe random 33 | tail -1; echo "== stress 31"; SP_PROBE_BIG=1 SP_PROBE_REPEAT=300 ./tc_probe random 31 | tail -1;
  for m in 2 3 1; do echo "== time 33, mode $m (2: BN prologue, no activation; 3: + LeakyReLU; 1: + ELU)"; SP_PROBE_ELU=$m ./tc_probe time 33; done;
  echo "== time 31 (bf16), BN + ELU"; SP_PROBE_ELU=1 ./tc_probe time 31; echo "== time 32 (generation 2), BN + ELU"; SP_PROBE_ELU=1 ./tc_probe time 32; echo "== time 32 (generation 2), BN, no activation"; SP_PROBE_ELU=2 ./tc_probe time 32;
  echo "== time 33 without TMA (plain-load staging), BN + ELU"; SP_TC3_NO_TMA=1 SP_PROBE_ELU=1 ./tc_probe time 33; echo "== time 31 without TMA, BN + ELU"; SP_TC3_NO_TMA=1 SP_PROBE_ELU=1 ./tc_probe time 31; } > ../../gpurun_out/r02/tc3_probe.log 2>&1
export SP_PROBE_ELU=1
./tc_probe time 33 > ../../gpurun_out/r02/ncu_plain_tc3.log 2>&1 && ncu --set full --clock-control none --import-source on -k regex:corr3_tc3 -s 3 -c 1 -o ../../gpurun_out/r02/ncu_tc3_fwd_elu ./tc_probe time 33 > ../../gpurun_out/r02/ncu_tc3.log 2>&1
unset SP_PROBE_ELU
cd ../..
timeout 600 python tools/diag_tc_layer.py 5,1,0 > gpurun_out/r02/diag_tc_layer.log 2>&1
timeout 900 python bench.py --steps 20 --warmup 3 --dump-breakdown gpurun_out/r02/kernel_breakdown.txt > gpurun_out/r02/bench_n1.json 2> gpurun_out/r02/bench_n1.err
timeout 600 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/r02/bench_reference_arm.json 2> gpurun_out/r02/bench_reference_arm.err
timeout 300 python bench.py --workload unet --steps 10 --warmup 3 --extras none --dump-breakdown gpurun_out/r02/kernel_breakdown_unet.txt > gpurun_out/r02/bench_unet_n1.json 2> gpurun_out/r02/bench_unet_n1.err
python bench.py --quick --steps 2 --warmup 1 --extras none > gpurun_out/r02/quick_plain.log 2>&1 && ncu --metrics gpu__time_duration.sum --clock-control none -c 700 --csv --log-file gpurun_out/r02/ncu_launches_bench_quick.csv python bench.py --quick --steps 2 --warmup 1 --extras none > gpurun_out/r02/ncu_launches.log 2>&1
tail -2 gpurun_out/r02/tc3_probe.log; tail -3 gpurun_out/r02/bench_n1.err; ls -la gpurun_out/r02
