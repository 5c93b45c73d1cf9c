#!/bin/bash
# One-GPU measurement session of a round: bench lines, launch list, ncu captures, config timings.
# Usage (on the GPU box, from the repo root): bash tools/round_measure.sh <round-tag>
# Every step runs under its own `timeout`: a hung kernel costs that step, not the box.
R=${1:-r2}
O=gpurun_out
T="timeout 240"
mkdir -p $O
$T python bench.py --impl reference --steps 3 --warmup 1 > $O/${R}_bench_reference_arm.json 2> $O/${R}_bench_reference_arm.err
$T python bench.py --steps 20 --warmup 3 > $O/${R}_bench_n1.json 2> $O/${R}_bench_n1.err
$T python tools/time_configs.py > $O/${R}_config_timings.txt 2>/dev/null
$T python tools/time_c3.py >> $O/${R}_config_timings.txt 2>/dev/null
$T python tools/time_gradmag.py >> $O/${R}_config_timings.txt 2>/dev/null
$T python tools/time_f32_1d.py >> $O/${R}_config_timings.txt 2>/dev/null
$T python tools/time_c5_parts.py >> $O/${R}_config_timings.txt 2>/dev/null
$T python tools/time_e2e.py 32 >> $O/${R}_config_timings.txt 2>/dev/null
for s in 0.5 1.0 1.5 2.0 2.5 3.0 4.0; do $T python tools/time_fused.py 512 $s reflect constant nearest mirror wrap >> $O/${R}_config_timings.txt 2>/dev/null; done
$T python tools/prof_corr.py >> $O/${R}_config_timings.txt 2>/dev/null
$T python tools/host_overhead.py 2>/dev/null | head -5 >> $O/${R}_config_timings.txt
$T python tools/time_at_size.py > $O/${R}_time_at_size.txt 2>/dev/null
$T python tools/clock_probe.py 512 2.0 reflect 2.0 > $O/${R}_clock_probe.txt 2>/dev/null
$T python tools/clock_probe.py 512 1.5 reflect 2.0 grad >> $O/${R}_clock_probe.txt 2>/dev/null
# launch list of the bench command (cold-cache, serialised: shares, not absolutes)
$T ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file $O/${R}_ncu_launch_list_bench.csv \
    python bench.py --steps 5 --warmup 3 --no-cpu --no-legs > /dev/null 2>&1
# full captures: the headline kernel, the single-launch gradient magnitude, C5's one-launch wide kernel, C3's streaming
# kernels, the dense 2-D correlate tile kernel
$T ncu --set full --clock-control none --import-source on -k regex:fws_kernel --launch-skip 2 -c 1 -f -o $O/${R}_prof_fused \
    python tools/prof_fused.py 512 2.0 reflect > /dev/null 2>&1
$T ncu --set full --clock-control none --import-source on -k regex:fws_kernel --launch-skip 2 -c 1 -f -o $O/${R}_prof_grad \
    python tools/prof_grad.py > /dev/null 2>&1
$T ncu --set full --clock-control none --import-source on -k regex:fws_kernel --launch-skip 2 -c 1 -f -o $O/${R}_prof_c5 \
    python tools/prof_c5.py > /dev/null 2>&1
$T ncu --set full --clock-control none --import-source on -k regex:exact_stream --launch-skip 4 -c 2 -f -o $O/${R}_prof_c3_stream \
    python tools/prof_c3.py > /dev/null 2>&1
$T ncu --set full --clock-control none --import-source on -k regex:correlate_2d_tile --launch-skip 4 -c 2 -f -o $O/${R}_prof_corr \
    python tools/prof_corr.py > /dev/null 2>&1
python tools/ncu_summary.py $O/${R}_prof_fused.ncu-rep > $O/${R}_fused_gaussian512_reflect_ncu_summary.txt 2>&1
python tools/ncu_summary.py $O/${R}_prof_grad.ncu-rep > $O/${R}_gradmag512_ncu_summary.txt 2>&1
python tools/ncu_summary.py $O/${R}_prof_c5.ncu-rep > $O/${R}_c5_wide_ncu_summary.txt 2>&1
python tools/ncu_summary.py $O/${R}_prof_c3_stream.ncu-rep > $O/${R}_c3_stream_ncu_summary.txt 2>&1
python tools/ncu_summary.py $O/${R}_prof_corr.ncu-rep > $O/${R}_correlate2d_tile_ncu_summary.txt 2>&1
tail -c 600 $O/${R}_bench_n1.json
