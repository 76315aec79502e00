"""Runs one HeatEquation::Problem verbosely through the host C API and lets the C++ layer print to stdout (captured by the
caller): the reference's per-step lines and ConvergenceTable (main.cc:689-719, 945-954, 1406-1411, 2045-2064, 3330-3370, 3464)."""
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from dealii_spirk_b200 import hostapi  # noqa: E402

if __name__ == "__main__":
    lib, tables, dim, params = sys.argv[1], sys.argv[2], int(sys.argv[3]), json.loads(sys.argv[4])
    host = hostapi.HostLib(lib, tables)
    with hostapi.Run(host, params, dim=dim, verbose=True) as run:
        run.run()
        print("TABLE_TEXT_BEGIN")
        print(run.table_text())
        print("DT=%r" % run.scalar("dt"))
