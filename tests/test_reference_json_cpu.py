"""The run configurations of the reference are read unchanged (north star: "the json/*.json run configs for ost/irk/spirk/
complex_spirk and the batched variants"): the key set AND the value typing of the reference's own files - numbers written
as strings ("IRKStages" : "5", "TimeStepSize" : "0.0"), the `DoRowMajor` / `Padding` / `UseSharedMemory` keys of
json/spirk_sm.json - parsed by the C++ host layer (main.cc:2970-3009).  The texts below are the reference's files with
NRefinements reduced so that the CPU double finishes in seconds; oracle comparison of the same runs: test_host_cpu.py."""
import json
import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "oracle"))
import host_checks as hc  # noqa: E402
import spirk_oracle as so  # noqa: E402
from dealii_spirk_b200 import hostapi  # noqa: E402

SCHEME_FILE = """{
    "FEDegree" : 1,
    "NRefinements" : 3,
    "TimeIntegrationScheme" : "%s",
    "IRKStages" : "5",
    "TimeStepSize" : "0.1",
    "EndTime" : "0.5",
    "OperatorType" : "MatrixFree",
    "BlockPreconditionerType" : "GMG",
    "InnerTolerance" : 0.0
}"""
SPIRK_SM = """{
    "FEDegree" : 1,
    "NRefinements" : 3,
    "TimeIntegrationScheme" : "spirk",
    "IRKStages" : "2",
    "TimeStepSize" : "0.0",
    "EndTime" : "0.5",
    "OperatorType" : "MatrixFree",
    "BlockPreconditionerType" : "GMG",
    "DoRowMajor" : true,
    "Padding" : 0,
    "UseSharedMemory" : true
}"""
OST = """{
    "FEDegree" : 4,
    "NRefinements" : 2,
    "TimeIntegrationScheme" : "ost",
    "IRKStages" : "3",
    "TimeStepSize" : "0.1",
    "EndTime" : "0.5"
}"""


@pytest.fixture(scope="module")
def cpu_host():
    return hostapi.HostLib(os.path.join(ROOT, "oracle", "_build", "libspirk_host_cpu.so"), hc.TABLES)


def run_text(host, text, dim=2):
    with hostapi.Run(host, json.loads(text), dim=dim) as run:  # (string-typed values stay strings through json.dumps)
        run.setup()
        while not run.finished():
            run.step()
        run.finish()
        return {"outer": run.array("outer_iterations"), "error_L2": run.array("error_L2"), "table": run.table_text()}


@pytest.mark.parametrize("scheme", ["irk", "irk_batched", "spirk", "complex_irk", "complex_irk_batched", "complex_spirk",
                                    "complex_spirk_batched"])
def test_reference_scheme_files(cpu_host, scheme):
    res = run_text(cpu_host, SCHEME_FILE % scheme)
    ora = so.run(scheme, 2, 1, 3, 5, 0.1, 0.5, outer_tol=1e-8)  # the keys the file omits take the defaults of main.cc:2943-3010
    assert len(res["outer"]) == 5
    assert np.allclose(res["error_L2"], np.array(ora["errors"])[:, 0], rtol=1e-6)
    cols = res["table"].strip().split("\n")[-2].split()
    vals = res["table"].strip().split("\n")[-1].split()
    assert int(vals[cols.index("n_stages")]) == 5 and int(vals[cols.index("fe_degree")]) == 1


def test_reference_shared_memory_file_with_automatic_time_step(cpu_host):
    res = run_text(cpu_host, SPIRK_SM)
    cols = res["table"].strip().split("\n")[-2].split()
    vals = res["table"].strip().split("\n")[-1].split()
    assert int(vals[cols.index("n_stages")]) == 2
    dt = float(vals[cols.index("dt")])
    assert 0.0 < dt < 0.5  # "TimeStepSize" : "0.0" selects the automatic rule of main.cc:3309-3318
    assert abs(float(vals[cols.index("final_t")]) - 0.5) < 1e-9 + dt


def test_reference_ost_file_needs_the_matrix_based_path(cpu_host):
    # json/ost.json carries no OperatorType / BlockPreconditionerType: the reference's defaults are MatrixBased + AMG
    # (Trilinos), which this build rejects with a clear message
    with pytest.raises(Exception) as e:
        run_text(cpu_host, OST)
    assert "MatrixBased" in str(e.value) or "MatrixFree" in str(e.value)


DEFAULT_JSON = """{
    "FEDegree" : 1,
    "NRefinements" : 3,
    "TimeIntegrationScheme" : "spirk",
    "IRKStages" : "4",
    "OuterTolerance" : "1e-12",
    "InnerTolerance" : "0.0",
    "TimeStepSize" : "0.1",
    "EndTime" : "1.0",
    "OperatorType" : "MatrixFree",
    "BlockPreconditionerType" : "GMG",
    "DoRowMajor" : true,
    "Padding" : -1,
    "DoOutputParaview" : false,
    "MaxRanks" : 4
}"""


def test_reference_scaling_template(cpu_host):
    """scripts/default.json, the template of the reference's scaling studies (every key its generator scripts set)"""
    res = run_text(cpu_host, DEFAULT_JSON)
    ora = so.run("spirk", 2, 1, 3, 4, 0.1, 1.0, outer_tol=1e-12)
    assert len(res["outer"]) == 10
    assert np.all(np.abs(res["outer"] - np.array(ora["integ"].n_outer)) <= 1)
    assert np.allclose(res["error_L2"], np.array(ora["errors"])[:, 0], rtol=1e-7)
