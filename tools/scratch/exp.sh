export CLB_PANEL_ROWS=256
timeout 280 compute-sanitizer --tool initcheck --print-limit 100000 python tools/scratch/san.py > /tmp/init.txt 2>&1
echo "device-side uninitialized reads: $(grep -c 'Uninitialized __global__ memory read' /tmp/init.txt)"; echo "host-side (cudaMemcpy source) reports: $(grep -c 'Host API memory access error' /tmp/init.txt)"; grep "parity\|ERROR SUMMARY" /tmp/init.txt
grep -A3 'Uninitialized __global__ memory read' /tmp/init.txt | grep " at \|Device Frame" | sort | uniq -c | sort -rn | head -8
timeout 200 compute-sanitizer --tool racecheck --print-limit 100 python tools/scratch/san.py 2>&1 | grep "Race reported\|and Read\|and Write\|RACECHECK SUMMARY" | sed 's/void clb:://; s/(const clb::Win.*int)//' | cut -c1-200 | sort | uniq -c | sort -rn | head -30
