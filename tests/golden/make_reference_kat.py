#!/usr/bin/env python3
"""Collect the known-answer vectors the JWave test-suite holds for the FWT/WPT path.

Runs only where /root/reference exists (the build container); the output
tests/golden/reference_kat.json is committed so nothing on the GPU box reads the reference.

Sources (relative to /root/reference):
  src/test/resources/testdata/haar_simple_input.txt, haar_level1_approx_manual.txt,
  haar_level1_detail_manual.txt      <- CrossValidationTest.testHaarTransformWithReference (:186-208)
  src/test/resources/testdata/filter_haar_{dec,rec}_{lo,hi}.txt
                                     <- CrossValidationTest.testHaarWaveletCoefficients (:158-181)
  src/test/resources/testdata/filter_db4_dec_{lo,hi}.txt  (PyWavelets 'db2' = JWave Daubechies2)
  src/test/resources/testdata/haar_constant_input.txt, haar_linear_input.txt
  src/test/java/jwave/GeneralTest.java:43                 <- the 8-sample round-trip vector
"""
import json
import os
import re

REF = "/root/reference"
TD = os.path.join(REF, "src/test/resources/testdata")
OUT = os.path.join(os.path.dirname(os.path.abspath(__file__)), "reference_kat.json")


def vec(name):
    with open(os.path.join(TD, name)) as fh:
        return [float(l) for l in fh if l.strip() and not l.startswith("#")]


def main():
    kat = {f[:-4]: vec(f) for f in sorted(os.listdir(TD)) if re.match(r"(haar_|filter_(haar|db))", f)}
    src = open(os.path.join(REF, "src/test/java/jwave/GeneralTest.java")).read()
    m = re.search(r"double\[\s*\]\s*arrTime\s*=\s*\{([^}]*)\}", src)
    kat["general_test_example"] = [float(t) for t in m.group(1).split(",")]
    with open(OUT, "w") as fh:
        json.dump(kat, fh, indent=1)
    print("wrote", OUT, sorted(kat))


if __name__ == "__main__":
    main()
