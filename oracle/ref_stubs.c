/*
 * ref_stubs.c - backend functions that the reference files compiled into oracle/_ref/libaocs_ref.so reference but the
 * oracle's drivers never reach (syscache lookups for InitSerTupInfo, record-type remapping, expanded-object and toast
 * flattening, foreign-server lookups).  Declared without the reference's headers so the signatures need not match;
 * each one reports itself and unwinds like an ereport(ERROR).  Test infrastructure.
 */
extern void ref_abort(const char *what);

#define REF_STUB(name) void name(void); void name(void) { ref_abort(#name); }
#ifndef REF_MOTION_LIB
REF_STUB(ReleaseSysCache)
REF_STUB(SearchSysCache1)	/* libmotion_ref.so answers pg_type look-ups from the driver's column list (ref_motion.c) */
#endif
REF_STUB(TRHandleTypeLists)
REF_STUB(build_tuple_node_list)
REF_STUB(deserializeNode)
REF_STUB(serializeNode)
REF_STUB(detoast_external_attr)
REF_STUB(format_type_be)
#ifndef REF_MOTION_LIB
REF_STUB(list_free_deep)	/* libmotion_ref.so compiles nodes/list.c itself */
#endif
REF_STUB(slot_getsomeattrs_int)
REF_STUB(DatumGetEOHP)
REF_STUB(EOH_flatten_into)
REF_STUB(EOH_get_flat_size)
REF_STUB(GetForeignServerSegByRelid)
REF_STUB(toast_flatten_tuple_to_datum)
/* reached only by the Motion layer's logging / by list.c entry points the drivers do not use (libmotion_ref.so) */
REF_STUB(GetConfigOption)
REF_STUB(GetConfigOptionResetString)
REF_STUB(appendStringInfo)
REF_STUB(initStringInfo)
REF_STUB(copyObjectImpl)
REF_STUB(equal)
REF_STUB(pg_qsort)
REF_STUB(qsort_arg)

/* data the same files reference */
struct { int dbid; int segindex; } GpIdentity = {0, 0};
