"""CPU tests of the small host-side tools that produce the committed profile summaries."""
import os
import subprocess
import sys

from conftest import ROOT

NCU_CSV = '''==PROF== Connected to process 1
"ID","Process ID","Process Name","Host Name","Kernel Name","Context","Stream","Block Size","Grid Size","Device","CC","Section Name","Metric Name","Metric Unit","Metric Value"
"0","1","python","h","void tr::keys_kernel(const float4 *, int)","1","13","(256, 1, 1)","(4096, 1, 1)","0","10.0","Command line profiler metrics","gpu__time_duration.sum","ns","12,000"
"1","1","python","h","tr::scan_kernel(const unsigned int *)","1","13","(256, 1, 1)","(1024, 1, 1)","0","10.0","Command line profiler metrics","gpu__time_duration.sum","us","6.0"
"2","1","python","h","void tr::keys_kernel(const float4 *, int)","1","13","(256, 1, 1)","(4096, 1, 1)","0","10.0","Command line profiler metrics","gpu__time_duration.sum","ns","10000"
'''


def test_summarise_launches(tmp_path):
    src = tmp_path / "launches.csv"
    src.write_text(NCU_CSV)
    r = subprocess.run([sys.executable, os.path.join(ROOT, "tools", "summarise_launches.py"), str(src), "cmd"],
                       capture_output=True, text=True, timeout=60)
    assert r.returncode == 0, r.stderr
    lines = [ln for ln in r.stdout.splitlines() if not ln.startswith("#")]
    assert lines[0] == "kernel,launches,total_ms,avg_us,share"
    assert lines[1].startswith("tr::keys_kernel,2,0.0220,11.00,0.78571")
    assert lines[2].startswith("tr::scan_kernel,1,0.0060,6.00,0.21429")
    assert "total 0.028 ms over 3 launches" in r.stdout
