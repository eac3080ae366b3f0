mkdir -p gpurun_out
O=gpurun_out/r02m_resident.txt; : > $O
python tools/probes/resident_probe.py >> $O 2>&1
LATE_ENV=ORB_B200_PROBE_SKIP_RESIZE=1 python tools/probes/resident_probe.py >> $O 2>&1
LATE_ENV=ORB_B200_PROBE_SKIP_RESIZE=2 python tools/probes/resident_probe.py >> $O 2>&1
cat $O
