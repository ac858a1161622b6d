/*
 * ref_plan_stubs.c - what oracle/_ref/libplan_ref.so (ref_plan.c: the f1 shim run on reference-built Plan trees) needs from
 * the backend besides the sources it compiles.  Test infrastructure.
 *
 * What the translate path really uses (the memory-context allocators under list.c / newNode, check_stack_depth, the hook
 * variables) is defined in ref_plan.c.  Everything here - executor entry points, the relation loader, the device library -
 * belongs to parts of the shim the harness never runs; calling one aborts through ref_exec_abort (a longjmp back into the test).
 */
extern void ref_exec_abort(const char *what);

/* declared without the reference's headers so the signatures need not match */
#define REF_STUB(name) void name(void); void name(void) { ref_exec_abort(#name); }
/* the executor / catalog side of the shim */
REF_STUB(ExecGetRangeTableRelation)
REF_STUB(ExecSetExecProcNode)
REF_STUB(ExecStoreVirtualTuple)
REF_STUB(MemoryContextRegisterResetCallback)
REF_STUB(RegisterXactCallback)
REF_STUB(standard_ExecutorStart)
REF_STUB(standard_ExecutorEnd)
REF_STUB(standard_planner)
REF_STUB(RegisterCustomScanMethods)
REF_STUB(DefineCustomStringVariable)
REF_STUB(getgpsegmentCount)
REF_STUB(get_opcode)
REF_STUB(get_rel_name)
REF_STUB(get_rel_type_id)
REF_STUB(get_typlenbyval)
REF_STUB(getTypeInputInfo)
REF_STUB(getTypeOutputInfo)
REF_STUB(get_promoted_array_type)
REF_STUB(get_relids_for_join)
REF_STUB(type_is_rowtype)
/* node support the reference's nodes / var.c reference from functions the harness does not reach */
REF_STUB(IncrementVarSublevelsUp)
REF_STUB(checkExprHasSubLink)
REF_STUB(copyObjectImpl)
REF_STUB(equal)
REF_STUB(makeString)
/* the shim's other files and the product's libraries (the harness stops at the translated plan) */
REF_STUB(cbgpu_shim_define_gucs)
REF_STUB(cbgpu_shim_dict)
REF_STUB(cbgpu_shim_interconnect)
REF_STUB(cbgpu_shim_load_relation)
REF_STUB(cbgpu_shim_qd_prepare)
REF_STUB(cb_CreateExecutorState)
REF_STUB(cb_ExecEndNode)
REF_STUB(cb_ExecInitNode)
REF_STUB(cb_ExecProcNode)
REF_STUB(cb_ExecReScan)
REF_STUB(cb_FreeExecutorState)
REF_STUB(cb_numeric_avg_serialize)
REF_STUB(cb_int8_avg_serialize)
REF_STUB(cb_estate_error)
REF_STUB(cb_slot_float8)
REF_STUB(cb_slot_int64)
REF_STUB(cb_slot_isnull)
REF_STUB(cb_slot_text)
REF_STUB(cbgpu_ctx_create)
REF_STUB(cbgpu_device_count)
REF_STUB(cbgpu_dict_lookup)
REF_STUB(cbgpu_rel_free)
