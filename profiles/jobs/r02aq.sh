# r02aq: probe — how much does the L1 size matter?  shared-memory carve-out forced to 100 % (228 KB: L1 28 KB) against the default (196 KB: L1 60 KB)
mkdir -p gpurun_out
for c in "" 100 86 ""; do
  echo "carveout '$c'" >> gpurun_out/r02aq_carveout.log
  if [ -n "$c" ]; then export RT_COOP_CARVEOUT=$c; else unset RT_COOP_CARVEOUT; fi
  python profiles/sweep_variants.py C3 8 0,0 2>&1 | grep variant >> gpurun_out/r02aq_carveout.log
  python profiles/sweep_variants.py C5 2 0 2>&1 | grep variant >> gpurun_out/r02aq_carveout.log
done
cat gpurun_out/r02aq_carveout.log
