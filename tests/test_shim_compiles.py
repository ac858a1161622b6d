"""The backend-side shim (integration/cbgpu_shim.c: ExecutorStart_hook, Plan -> CbPlan translation, the replaced
ExecProcNode) type-checks against the REFERENCE's own headers (not gpu; only where /root/reference exists).

This is the drop-in boundary of SURVEY.md 8b written as compilable code: every reference struct member, enum and
function it touches (PlanState.ExecProcNode, HashJoin.hashkeys, Agg.grpColIdx, Motion.hashExprs, ExecStoreVirtualTuple,
MemoryContextRegisterResetCallback ...) must exist with the type the shim assumes, and every cb_* entry point it
calls must match include/cb_exec.h.  Generated headers (catalog/*_d.h, errcodes.h, fmgrprotos.h, pg_config.h) are the
stand-ins under oracle/ref_shim/."""
import os
import subprocess

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
REF = "/root/reference/src/include"
SOURCES = ["cbgpu_shim.c", "cbgpu_shim_storage.c", "cbgpu_shim_motion.c", "cbgpu_shim_interconnect.c"]


@pytest.mark.skipif(not os.path.isdir(REF), reason="reference headers only in the build container")
@pytest.mark.parametrize("source", SOURCES)
def test_shim_type_checks_against_reference_headers(source):
    """cbgpu_shim.c: the operator boundary; cbgpu_shim_storage.c: catalogs (pg_aocsseg, pg_attribute_encoding, pg_aovisimap,
    relation options) -> cb_aocs_load_segfile; cbgpu_shim_motion.c: executor rows <-> the configured MotionIPCLayer's
    tuple chunks (SendTupleChunkToAMS / RecvTupleChunkFromAny) through cb_tupser_*; cbgpu_shim_interconnect.c: the session's NCCL
    communicator, its rendezvous token shipped from the QD as a synced GUC"""
    p = subprocess.run(["gcc", "-fsyntax-only", "-Wall", "-Werror", "-Wno-unused-function",
                        "-I" + os.path.join(ROOT, "oracle", "ref_shim"), "-I" + REF, "-I" + os.path.join(ROOT, "include"),
                        os.path.join(ROOT, "integration", source)], capture_output=True, text=True)
    assert p.returncode == 0, p.stderr[-4000:]


@pytest.mark.skipif(not os.path.isdir(REF), reason="reference headers only in the build container")
def test_shim_module_links_against_the_libraries(tmp_path):
    """the four files as ONE loadable module (what shared_preload_libraries would load): it exports PG_MODULE_MAGIC's
    Pg_magic_func and _PG_init, leaves no cbgpu_shim_* symbol of its own undefined, and every cb_* / cbgpu_* symbol it
    needs is exported by libcbexec.so / libcbgpu.so - everything else it leaves undefined is the backend's"""
    so = str(tmp_path / "cbgpu_shim.so")
    p = subprocess.run(["gcc", "-shared", "-fPIC", "-Wall", "-Werror", "-Wno-unused-function", "-I" + os.path.join(ROOT, "oracle", "ref_shim"),
                        "-I" + REF, "-I" + os.path.join(ROOT, "include")] + [os.path.join(ROOT, "integration", f) for f in SOURCES] +
                       ["-o", so], capture_output=True, text=True)
    assert p.returncode == 0, p.stderr[-4000:]

    def syms(path, flag):
        out = subprocess.run(["nm", "-D", flag, path], capture_output=True, text=True).stdout
        return {ln.split()[-1].split("@")[0] for ln in out.splitlines() if ln.strip()}
    defined = syms(so, "--defined-only")
    undefined = syms(so, "--undefined-only")
    assert {"Pg_magic_func", "_PG_init"} <= defined
    ours = {s for s in undefined if s.startswith("cb_") or s.startswith("cbgpu_")}
    assert not {s for s in ours if s.startswith("cbgpu_shim_")}, ours
    exported = syms(os.path.join(ROOT, "cloudberry_b200", "libcbexec.so"), "--defined-only") | \
        syms(os.path.join(ROOT, "cloudberry_b200", "libcbgpu.so"), "--defined-only")
    assert ours and ours <= exported, ours - exported
    # what is left is the backend's: spot-check the ones the boundary stands on
    assert {"ExecutorStart_hook", "standard_ExecutorStart", "ExecStoreVirtualTuple", "MemoryContextRegisterResetCallback",
            "DefineCustomStringVariable", "CdbDispatchSetCommand", "CurrentMotionIPCLayer"} <= undefined
