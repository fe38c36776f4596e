"""Invariants of the CUDA sources that no run-time test would catch.

Programmatic dependent launch (csrc/kernels.h): every launch carries the stream-serialization attribute, so a kernel may be
set up before its predecessor in the stream has finished; what keeps the order of memory effects is that EVERY kernel's first
statement is pdl_sync() (griddepcontrol.wait). A kernel without it, or a plain <<<>>> launch next to attributed ones, would be
a silent race."""
import glob
import os
import re

CSRC = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "baofit_b200", "csrc")


def _sources():
    return sorted(glob.glob(os.path.join(CSRC, "*.cu")) + glob.glob(os.path.join(CSRC, "*.cuh")))


def test_every_kernel_starts_with_the_dependency_wait():
    kernels = 0
    for path in _sources():
        lines = open(path).read().split("\n")
        i = 0
        while i < len(lines):
            if "__global__" in lines[i]:
                j = i
                while not lines[j].rstrip().endswith("{"):  # end of the (possibly multi-line) signature
                    j += 1
                body = j + 1
                while lines[body].strip() == "" or lines[body].strip().startswith("//"):
                    body += 1
                assert lines[body].strip() == "pdl_sync();", "%s:%d: kernel does not start with pdl_sync()" % (path, i + 1)
                kernels += 1
                i = j
            i += 1
    assert kernels >= 40


def test_all_launches_go_through_the_launch_helper():
    for path in _sources():
        text = open(path).read()
        assert "<<<" not in text, "%s: plain <<<>>> launch (use bf::Launch so the kernel chain keeps its programmatic edges)" % path
    helper = open(os.path.join(CSRC, "kernels.h")).read()
    assert "cudaLaunchAttributeProgrammaticStreamSerialization" in helper
    assert re.search(r"pdl_sync\(\)\s*\{\s*asm volatile\(\"griddepcontrol\.wait;\"", helper)
    assert "griddepcontrol.launch_dependents" not in helper.split("#ifdef __CUDACC__")[1].split("#endif")[0]  # no early trigger
