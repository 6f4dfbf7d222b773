# Round-2 measurement pass (run through gpurun from the repo root): probes, bench lines, launch list, ncu captures.
set -x
O=gpurun_out/r02g
mkdir -p $O
cd tools/tc_probe
{ for g in 0 3 1; do echo "== check4 (generation 2), grid cap $g"; SP_WTC4_GRID=$g ./wgrad_probe check4 6; done;
  echo "== check4 drain 2"; ./wgrad_probe check4 2; echo "== check (generation 1)"; ./wgrad_probe check 2 64; echo "== check24 (sp_wgrad_tc24)"; ./wgrad_probe check24 2;
  for de in 4 6 8; do echo "== time4 drain $de"; ./wgrad_probe time4 $de; done; echo "== time generation 1"; ./wgrad_probe time 2 64;
  echo "== check244 (24 -> 20 channels: four slice pairs in one launch)"; ./wgrad_probe check244 6; echo "== time244 (24 -> 24, batch 32, 14x58x58)"; ./wgrad_probe time244 6; echo "== time24 (sp_wgrad_tc24)"; ./wgrad_probe time24 2;
  for g in 0 3; do echo "== checks2 (stride 2, 16 -> 24), grid cap $g"; SP_WTC4_GRID=$g ./wgrad_probe checks2 12; done; echo "== times2 (Cae3D.py:48 at batch 24)"; ./wgrad_probe times2 12; } > ../../$O/wgrad_tc4_probe.log 2>&1
{ echo "== random 33"; SP_PROBE_BIG=1 ./tc_probe random 33; echo "== stress 33"; SP_PROBE_BIG=1 SP_PROBE_REPEAT=200 ./tc_probe random 33 | tail -1;
  for m in 2 3 1; do echo "== time 33, mode $m (2: BN prologue, no activation; 3: + LeakyReLU; 1: + ELU)"; SP_PROBE_ELU=$m ./tc_probe time 33; done;
  echo "== time 31 (bf16), BN + ELU"; SP_PROBE_ELU=1 ./tc_probe time 31; } > ../../$O/tc3_probe.log 2>&1
./wgrad_probe time4 6 > ../../$O/ncu_plain_wgrad_tc4.log 2>&1 && ncu --set full --clock-control none --import-source on -k regex:wgrad3_tc4 -s 2 -c 1 -o ../../$O/ncu_wgrad_tc4 ./wgrad_probe time4 6 > ../../$O/ncu_wgrad_tc4.log 2>&1
cd ../..
timeout 900 python bench.py --steps 20 --warmup 3 --dump-breakdown $O/kernel_breakdown.txt > $O/bench_n1.json 2> $O/bench_n1.err
timeout 600 python bench.py --impl reference --steps 3 --warmup 1 > $O/bench_reference_arm.json 2> $O/bench_reference_arm.err
timeout 300 python bench.py --workload unet --steps 10 --warmup 3 --extras none --dump-breakdown $O/kernel_breakdown_unet.txt > $O/bench_unet_n1.json 2> $O/bench_unet_n1.err
python bench.py --quick --steps 2 --warmup 1 --extras none > $O/quick_plain.log 2>&1 && ncu --metrics gpu__time_duration.sum --clock-control none -c 700 --csv --log-file $O/ncu_launches_bench_quick.csv python bench.py --quick --steps 2 --warmup 1 --extras none > $O/ncu_launches.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:corr3_tc3_kernel -s 4 -c 2 -o $O/ncu_tc3_step python bench.py --quick --steps 1 --warmup 1 --extras none > $O/ncu_tc3_step.log 2>&1
timeout 300 python tools/time_e2e_parts.py > $O/e2e_parts.log 2>&1
tail -3 $O/bench_n1.err; ls -la $O
