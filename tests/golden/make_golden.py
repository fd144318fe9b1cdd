#!/usr/bin/env python3
"""Generates tests/golden/*.npz from the reference itself.  Runs only in the build container
(/root/reference is not present on the GPU box); the outputs are committed.

1. ewald_table_xi*.npz — the real-space table of Stokes::setParams.  The loop body of
   PSEv1/Stokes.cc (from `double r = double( kk ) * dr + dr;` to `// Save values to table`) is
   read from the reference source where it lies, wrapped in a tiny C++ main under
   oracle/_ref/ (git-ignored) and compiled with g++; nothing from the reference is copied into
   the repository.
2. shear_functions.npz — values of the reference's header-only shear classes
   (PSEv1/SpecificShearFunction.h, compiled against a two-line pybind11 stub).
"""
import os
import subprocess
import sys

import numpy as np

REF = "/root/reference/PSEv1"
HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
OUT = os.path.join(ROOT, "oracle", "_ref")


def build_table_program():
    src = open(os.path.join(REF, "Stokes.cc")).read().split("\n")
    start = next(i for i, l in enumerate(src) if "double r = double( kk ) * dr + dr;" in l)
    end = next(i for i, l in enumerate(src) if "// Save values to table" in l)
    body = "\n".join(src[start:end])
    prog = r"""
#include <cmath>
#include <cstdio>
#include <cstdlib>
using namespace std;
int main(int argc, char** argv) {
    double xi = atof(argv[1]); int nR = atoi(argv[2]);
    double dr = 0.0010000000000000; double Pi = 3.141592653589793; double a = 1.0;
    for (int kk = 0; kk < nR; kk++) {
""" + body + r"""
        float fI = (float)Imrr, fr = (float)rr;
        printf("%.9g %.9g %.17g %.17g\n", fI, fr, Imrr, rr);
    }
    return 0;
}
"""
    os.makedirs(OUT, exist_ok=True)
    cpp = os.path.join(OUT, "gen_table.cpp")
    exe = os.path.join(OUT, "gen_table")
    open(cpp, "w").write(prog)
    subprocess.check_call(["g++", "-O1", "-ffp-contract=off", "-o", exe, cpp])
    return exe


def gen_tables():
    exe = build_table_program()
    for xi in (0.5, 0.3, 0.8):
        err = 1e-3
        rcut = np.float32(np.sqrt(-np.log(np.float32(err)), dtype=np.float32)) / np.float32(xi)
        n = int(np.float32(rcut) / np.float32(0.001) - np.float32(1))  # PSEv1/Stokes.cc:310
        out = subprocess.check_output([exe, repr(float(np.float32(xi))), str(n + 1)]).decode().split("\n")
        rows = np.array([[float(x) for x in l.split()] for l in out if l.strip()])
        np.savez_compressed(os.path.join(HERE, f"ewald_table_xi{xi}.npz"), xi=np.float32(xi), ewald_n=n,
                            fg32=rows[:, :2].astype(np.float32), fg64=rows[:, 2:])
        print("table xi", xi, "entries", len(rows), "f(2.0), g(2.0) =", rows[1999, 2:] if len(rows) > 1999 else None)


def gen_shear():
    stub = os.path.join(OUT, "stub", "hoomd", "extern", "pybind", "include", "pybind11")
    os.makedirs(stub, exist_ok=True)
    open(os.path.join(stub, "pybind11.h"), "w").write("#pragma once\n#include <memory>\nnamespace pybind11 { class module; }\n")
    prog = r"""
#include <memory>
#include <cstdio>
#include "SpecificShearFunction.h"
int main() {
    const double dt = 1e-3;
    std::shared_ptr<ShearFunction> f[6];
    f[0] = std::make_shared<SteadyShearFunction>(1.5, 10u, dt);
    f[1] = std::make_shared<SinShearFunction>(2.0, 3.0, 10u, dt);
    f[2] = std::make_shared<ChirpShearFunction>(0.1, 1.0, 50.0, 2.0, 10u, dt);
    f[3] = std::make_shared<TukeyWindowFunction>(2.0, 0.5, 10u, dt);
    f[4] = std::make_shared<WindowedFunction>(f[2], f[3]);
    f[5] = std::make_shared<ShearFunction>();
    for (int k = 0; k < 6; ++k)
        for (unsigned t = 0; t <= 2400; t += 37)
            printf("%d %u %.17g %.17g %u\n", k, t, f[k]->getShearRate(t), f[k]->getStrain(t), f[k]->getOffset());
    return 0;
}
"""
    cpp = os.path.join(OUT, "gen_shear.cpp")
    exe = os.path.join(OUT, "gen_shear")
    open(cpp, "w").write(prog)
    subprocess.check_call(["g++", "-O1", "-I", REF, "-I", os.path.join(OUT, "stub"), "-o", exe, cpp])
    rows = np.array([[float(x) for x in l.split()] for l in subprocess.check_output([exe]).decode().split("\n") if l.strip()])
    np.savez_compressed(os.path.join(HERE, "shear_functions.npz"), rows=rows)
    print("shear rows", rows.shape)


if __name__ == "__main__":
    if not os.path.isdir(REF):
        sys.exit("reference tree not present; golden files are generated in the build container only")
    gen_tables()
    gen_shear()
