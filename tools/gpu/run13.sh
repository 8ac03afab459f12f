python tools/sanitize_probe.py > gpurun_out/sanitize_plain.log 2>&1 && compute-sanitizer --tool racecheck --racecheck-report all python tools/sanitize_probe.py --scene-only > gpurun_out/racecheck_r2.log 2>&1
tail -25 gpurun_out/racecheck_r2.log
