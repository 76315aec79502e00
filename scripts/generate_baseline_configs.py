#!/usr/bin/env python3
"""Write the JSON run configurations of BASELINE.json into json/ (the Python side only generates
parameter files, as the reference's scripts/*.py do, e.g. scripts/large_scaling.py:5-18).

Key set and defaults are the reference's (main.cc:2970-3009, scripts/default.json).  The reference's
shipped json/*.json use FEDegree 1 / IRKStages 5; the BASELINE configs are the
Q4 variants generated here.
"""
import json
import os

OUT = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "json")
DEFAULT = {"FEDegree": 1, "NRefinements": 8, "TimeIntegrationScheme": "spirk", "IRKStages": 4, "OuterTolerance": 1e-12,
           "InnerTolerance": 0.0, "TimeStepSize": 0.1, "EndTime": 1.0, "OperatorType": "MatrixFree",
           "BlockPreconditionerType": "GMG", "DoRowMajor": True, "Padding": -1, "DoOutputParaview": False}


def write(name, **over):
    cfg = dict(DEFAULT)
    cfg.update(over)
    with open(os.path.join(OUT, name), "w") as f:
        json.dump(cfg, f, indent=4)
        f.write("\n")


def main():
    os.makedirs(OUT, exist_ok=True)
    # config 1: 2-D Q2, 5 refinements, one-step-theta, GMG-preconditioned CG (run with --dim 2)
    write("ost.json", FEDegree=2, NRefinements=5, TimeIntegrationScheme="ost", IRKStages=3, EndTime=0.5, OuterTolerance=1e-8)
    # config 2: 3-D Q4, IRK q=2, single GPU
    write("irk.json", FEDegree=4, NRefinements=6, TimeIntegrationScheme="irk", IRKStages=2, EndTime=0.5, OuterTolerance=1e-8)
    write("irk_batched.json", FEDegree=4, NRefinements=6, TimeIntegrationScheme="irk_batched", IRKStages=2, EndTime=0.5,
          OuterTolerance=1e-8)
    # config 3: 3-D Q4, SPIRK q=4, one stage per GPU on 4 GPUs
    write("spirk.json", FEDegree=4, NRefinements=6, TimeIntegrationScheme="spirk", IRKStages=4, EndTime=0.5, OuterTolerance=1e-8)
    # config 4: 3-D Q4, complex SPIRK q=8 batched, conjugate stage pairs across the GPUs
    write("complex_spirk_batched.json", FEDegree=4, NRefinements=6, TimeIntegrationScheme="complex_spirk_batched", IRKStages=8,
          EndTime=0.5, OuterTolerance=1e-8)
    write("complex_spirk.json", FEDegree=4, NRefinements=6, TimeIntegrationScheme="complex_spirk", IRKStages=8, EndTime=0.5,
          OuterTolerance=1e-8)
    write("complex_irk.json", FEDegree=4, NRefinements=5, TimeIntegrationScheme="complex_irk", IRKStages=4, EndTime=0.5,
          OuterTolerance=1e-8)
    write("complex_irk_batched.json", FEDegree=4, NRefinements=5, TimeIntegrationScheme="complex_irk_batched", IRKStages=4,
          EndTime=0.5, OuterTolerance=1e-8)
    # the reference's json/spirk_sm.json (shared-memory mixing, automatic time step) at the bench size
    write("spirk_sm.json", FEDegree=4, NRefinements=6, TimeIntegrationScheme="spirk", IRKStages=2, TimeStepSize=0.0, EndTime=0.5,
          OuterTolerance=1e-8, Padding=0, UseSharedMemory=True)
    # config 5: large-scaling SPIRK q=8 Q4 r=7 (135 005 697 DoFs x 8 stages), scripts/default.json parameters
    write("spirk_large.json", FEDegree=4, NRefinements=7, TimeIntegrationScheme="spirk", IRKStages=8)


if __name__ == "__main__":
    main()
