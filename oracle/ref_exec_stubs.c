/*
 * ref_exec_stubs.c - backend functions the reference files compiled into oracle/_ref/libexec_ref.so link against but
 * oracle/ref_exec.c's drivers never reach (catalog lookups of makeCdbHashForRelation, the binary send/receive functions,
 * HyperLogLog, set-returning helpers, locale-aware comparison).  Declared without the reference's headers so the
 * signatures need not match; each one reports itself and unwinds like an ereport(ERROR).  Test infrastructure.
 */
extern void ref_exec_abort(const char *what);

#define REF_STUB(name) void name(void); void name(void) { ref_exec_abort(#name); }
/* libpq/pqformat.c and common/stringinfo.c are compiled in (the aggregate serialisation functions write their bytea through
 * them); what THEY reference but this library never reaches: */
REF_STUB(pg_server_to_client)
REF_STUB(pg_client_to_server)
REF_STUB(pvsnprintf)
void	   *PqCommMethods = 0;
REF_STUB(ArrayGetIntegerTypmods)
REF_STUB(BackoffBackendTickExpired)
REF_STUB(GetDefaultOpClass)
REF_STUB(IsBinaryCoercible)
REF_STUB(MemoryContextStats)
REF_STUB(ProcessInterrupts)
REF_STUB(ReleaseCatCacheList)
REF_STUB(ResolveOpClass)
REF_STUB(SearchSysCacheList)
REF_STUB(UpdateTimeAtomically)
REF_STUB(addHyperLogLog)
REF_STUB(cdblegacyhash_null)
REF_STUB(cstring_to_text)
REF_STUB(cstring_to_text_with_len)
REF_STUB(end_MultiFuncCall)
REF_STUB(estimateHyperLogLog)
#ifndef REF_PLAN_LIB				/* libplan_ref.so links the reference's nodes/nodeFuncs.c, which defines these */
REF_STUB(exprTypmod)
#endif
REF_STUB(float4in)
REF_STUB(float8in)
REF_STUB(format_type_be)
REF_STUB(get_legacy_cdbhash_opclass_for_base_type)
REF_STUB(get_opclass_family)
REF_STUB(get_opfamily_member)
REF_STUB(get_opfamily_proc)
REF_STUB(initHyperLogLog)
REF_STUB(init_MultiFuncCall)
REF_STUB(lookup_type_cache)
REF_STUB(per_MultiFuncCall)
REF_STUB(pg_database_encoding_max_length)
REF_STUB(pg_mbcharcliplen)
REF_STUB(pg_mbcliplen)
REF_STUB(pg_mbstrlen_with_len)
REF_STUB(pg_newlocale_from_collation)
#ifndef REF_PLAN_LIB
REF_STUB(relabel_to_typmod)
#endif
REF_STUB(text_to_cstring)
REF_STUB(textsend)
REF_STUB(toast_raw_datum_size)
REF_STUB(varstr_cmp)
REF_STUB(varstr_sortsupport)
REF_STUB(write_stderr)
REF_STUB(pg_detoast_datum_copy)
