#!/usr/bin/env python
"""Build the UNMODIFIED reference engine (Genomics-HSE/VGsim, Cython) out of tree.

TEST INFRASTRUCTURE ONLY.  Nothing in the product path (vgsim_b200/) imports what
this script produces.  The reference sources are read where they lie under
/root/reference (never copied into the repository): they are staged in a scratch
directory under /tmp, cythonized there, and only the BUILT artefacts land in
oracle/_ref/ (git-ignored, but shipped to the GPU box by gpurun):

    oracle/_ref/VGsim/_BirthDeath*.so      the reference engine (src/_BirthDeath.pyx + *.pxi)
    oracle/_ref/VGsim/{_interface,IO}.pyc.bin  the reference's Python wrapper and writers, byte-compiled (sourceless)
    oracle/_ref/mc_lib/rndm*.so            shim for the un-vendored third-party dependency
    oracle/_ref/{prettytable,tskit,matplotlib}   import stubs (diagnostics only, never called)

`mc_lib` (pinned v0.4.1 in the reference's pyproject.toml:9,33) is NOT in this image
and there is no network; the shim restates its published behaviour:
RndmWrapper(seed=(entropy, num)) wraps numpy PCG64(SeedSequence(entropy, spawn_key=(num,)))
and uniform() is bitgen.next_double.  The seeding rule cannot be
checked offline against upstream mc_lib -> seed-for-seed parity with upstream builds is
"unpinned"; parity with THIS build is what tests/golden/ pins.

Cython 3 note: legacy_implicit_noexcept=True gives the reference its
fast build, so the CPU baseline is not handicapped.
"""
import os, shutil, subprocess, sys, tempfile, textwrap, glob

REF = os.environ.get("VGSIM_REFERENCE", "/root/reference")
HERE = os.path.dirname(os.path.abspath(__file__))
OUT = os.path.join(HERE, "_ref")


def main():
    if not os.path.isdir(os.path.join(REF, "src")):
        print("build_ref: %s not present; keeping any prebuilt oracle/_ref" % REF)
        return 0
    work = tempfile.mkdtemp(prefix="vgsim_ref_build_")
    pkg = os.path.join(work, "VGsim")
    os.makedirs(pkg)
    for f in ("_BirthDeath.pyx", "fast_choose.pxi", "events.pxi", "models.pxi"):
        shutil.copy(os.path.join(REF, "src", f), pkg)
    # the engine is driven directly (BirthDeathModel).  The reference's pure-Python wrapper and writers
    # (src/_interface.py, src/IO.py) are byte-compiled below into sourceless .pyc files: the parity tests run the
    # reference's own Simulator class over the new engine (drop-in proof) and its own Newick / TSV writers
    # (writer parity) on the GPU box, where /root/reference does not exist.
    open(os.path.join(pkg, "__init__.py"), "w").write(
        "from ._BirthDeath import BirthDeathModel\n")
    mc = os.path.join(work, "mc_lib")
    os.makedirs(mc)
    open(os.path.join(mc, "__init__.py"), "w").write("")
    open(os.path.join(mc, "rndm.pxd"), "w").write(textwrap.dedent("""\
        from numpy.random cimport bitgen_t
        cdef class RndmWrapper():
            cdef bitgen_t *rng
            cdef object py_gen
            cdef inline double uniform(self) noexcept nogil:
                return self.rng.next_double(self.rng.state)
        """))
    open(os.path.join(mc, "rndm.pyx"), "w").write(textwrap.dedent("""\
        # cython: language_level=3
        from cpython.pycapsule cimport PyCapsule_GetPointer
        from numpy.random cimport bitgen_t
        from numpy.random import PCG64, SeedSequence
        cdef class RndmWrapper():
            def __init__(self, seed=(1234, 0), bitgen_kind=None):
                entropy, num = seed
                self.py_gen = PCG64(SeedSequence(entropy, spawn_key=(num,)))
                self.rng = <bitgen_t *>PyCapsule_GetPointer(self.py_gen.capsule, "BitGenerator")
        """))
    open(os.path.join(work, "setup.py"), "w").write(textwrap.dedent("""\
        import os, numpy
        from setuptools import setup, Extension
        from Cython.Build import cythonize
        npd = os.path.dirname(numpy.__file__)
        exts = [
            Extension("mc_lib.rndm", ["mc_lib/rndm.pyx"], include_dirs=[numpy.get_include()]),
            Extension("VGsim._BirthDeath", ["VGsim/_BirthDeath.pyx"], language="c++",
                      include_dirs=[numpy.get_include()],
                      library_dirs=[os.path.join(npd, "random", "lib"), os.path.join(npd, "_core", "lib")],
                      libraries=["npyrandom", "npymath"],
                      extra_compile_args=["-O3", "-march=x86-64-v3", "-ffp-contract=off", "-w"]),
        ]
        setup(name="vgsim_ref", ext_modules=cythonize(exts, language_level=3, include_path=["."],
              compiler_directives={"legacy_implicit_noexcept": True}))
        """))
    # stubs needed at import time only
    stubs = os.path.join(work, "stubs")
    os.makedirs(os.path.join(stubs, "matplotlib"))
    open(os.path.join(stubs, "prettytable.py"), "w").write(textwrap.dedent("""\
        class PrettyTable:
            def __init__(self, *a, **k):
                self.field_names = []
                self.rows = []
            def add_row(self, row):
                self.rows.append(list(row))
            def __str__(self):
                return "\\n".join(["\\t".join(map(str, self.field_names))] +
                                 ["\\t".join(map(str, r)) for r in self.rows])
        """))
    open(os.path.join(stubs, "tskit.py"), "w").write("")
    open(os.path.join(stubs, "matplotlib", "__init__.py"), "w").write("")
    open(os.path.join(stubs, "matplotlib", "pyplot.py"), "w").write("")
    env = dict(os.environ)
    env["PYTHONPATH"] = stubs + os.pathsep + env.get("PYTHONPATH", "")
    subprocess.check_call([sys.executable, "setup.py", "-q", "build_ext", "--inplace"], cwd=work, env=env)
    # publish built artefacts only
    if os.path.isdir(OUT):
        shutil.rmtree(OUT)
    os.makedirs(os.path.join(OUT, "VGsim"))
    os.makedirs(os.path.join(OUT, "mc_lib"))
    for so in glob.glob(os.path.join(pkg, "*.so")):
        shutil.copy(so, os.path.join(OUT, "VGsim"))
    shutil.copy(os.path.join(pkg, "__init__.py"), os.path.join(OUT, "VGsim"))
    import py_compile
    for f in ("_interface.py", "IO.py"):
        # stored as *.pyc.bin: file-sync tools (the GPU-box snapshot among them) tend to drop *.pyc
        py_compile.compile(os.path.join(REF, "src", f), cfile=os.path.join(OUT, "VGsim", f + "c.bin"), dfile="VGsim/" + f,
                           doraise=True)
    for so in glob.glob(os.path.join(mc, "*.so")):
        shutil.copy(so, os.path.join(OUT, "mc_lib"))
    shutil.copy(os.path.join(mc, "__init__.py"), os.path.join(OUT, "mc_lib"))
    shutil.copy(os.path.join(stubs, "prettytable.py"), OUT)
    shutil.copy(os.path.join(stubs, "tskit.py"), OUT)
    shutil.copytree(os.path.join(stubs, "matplotlib"), os.path.join(OUT, "matplotlib"))
    shutil.rmtree(work, ignore_errors=True)
    print("build_ref: reference engine built into", OUT)
    return 0


if __name__ == "__main__":
    sys.exit(main())
