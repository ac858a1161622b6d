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


@pytest.mark.skipif(not os.path.isdir(REF), reason="reference headers only in the build container")
@pytest.mark.parametrize("source", ["cbgpu_shim.c", "cbgpu_shim_storage.c", "cbgpu_shim_motion.c"])
def test_shim_type_checks_against_reference_headers(source):
    """cbgpu_shim.c: the operator boundary; cbgpu_shim_storage.c: catalogs (pg_aocsseg, pg_attribute_encoding, pg_aovisimap,
    relation options) -> cb_aocs_load_segfile; cbgpu_shim_motion.c: executor rows <-> the configured MotionIPCLayer's
    tuple chunks (SendTupleChunkToAMS / RecvTupleChunkFromAny) through cb_tupser_*"""
    p = subprocess.run(["gcc", "-fsyntax-only", "-Wall", "-Werror", "-Wno-unused-function",
                        "-I" + os.path.join(ROOT, "oracle", "ref_shim"), "-I" + REF, "-I" + os.path.join(ROOT, "include"),
                        os.path.join(ROOT, "integration", source)], capture_output=True, text=True)
    assert p.returncode == 0, p.stderr[-4000:]
