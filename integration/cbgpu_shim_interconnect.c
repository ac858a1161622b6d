/*
 * cbgpu_shim_interconnect.c - the NCCL communicator behind CbEState.es_interconnect inside a Cloudberry session.
 *
 * cbgpu_shim.c asks for it once per query (cbgpu_shim_interconnect).  A communicator needs one 128-byte rendezvous token,
 * made by one participant and known to all (cbgpu_motion_unique_id / cbgpu_motion_create, include/cbgpu.h:326-327); the
 * participants are the segments' QE backends, rank = GpIdentity.segindex of getgpsegmentCount().  The token travels the way
 * the reference already moves per-session settings from the QD to its QEs, without patching core:
 *
 *   - a custom string GUC, cbgpu.nccl_id (hex), flagged GUC_GPDB_NEED_SYNC (utils/guc.h:247): gangs created later receive it
 *     in their connection options (cdbgang.c:455-470 makeOptions -> "-c name=value");
 *   - gangs that exist already get it through a dispatched SET (CdbDispatchSetCommand, cdb/cdbdisp_query.h:69), issued by the
 *     QD's ExecutorStart hook before the plan is dispatched (cbgpu_shim_qd_prepare, called from cbgpu_shim.c's hook on the QD).
 *
 * On a QE the communicator is created lazily by the first query that needs it and kept for the session (creating one is a
 * collective over all segments: every QE of the slice reaches this point for the same query); a new token - the QD starts a
 * new one when a query ended in error, since an interrupted collective leaves a communicator unusable - replaces it.
 * Only the QE process that owns the segment's device may do this (the GPU sub-tree lives in one slice).
 *
 * Type-checked against the reference's headers (tests/test_shim_compiles.py) and link-checked against libcbexec / libcbgpu
 * (every cb_* / cbgpu_* symbol the shim module needs is exported by them); not run here (no backend).
 */
#include "postgres.h"

#include "cdb/cdbdisp_query.h"
#include "cdb/cdbutil.h"
#include "cdb/cdbvars.h"
#include "nodes/execnodes.h"
#include "utils/builtins.h"
#include "utils/guc.h"
#include "utils/memutils.h"

#include "cb_exec.h"

#define NCCL_ID_BYTES 128

static char *cbgpu_nccl_id_hex = NULL;	/* the GUC's value: 256 hex digits, or empty */
static char	current_id_hex[2 * NCCL_ID_BYTES + 1];	/* token of the communicator below */
static cbgpu_motion *session_motion = NULL;
static CbInterconnect *session_ic = NULL;

void		cbgpu_shim_define_gucs(void);
void		cbgpu_shim_qd_prepare(bool previous_query_failed);
CbInterconnect *cbgpu_shim_interconnect(cbgpu_ctx *ctx, EState *estate);

static bool
check_nccl_id(char **newval, void **extra, GucSource source)
{
	const char *v = *newval;

	(void) extra;
	(void) source;
	if (v == NULL || v[0] == '\0')
		return true;
	if (strlen(v) != 2 * NCCL_ID_BYTES || strspn(v, "0123456789abcdef") != 2 * NCCL_ID_BYTES)
	{
		GUC_check_errdetail("cbgpu.nccl_id must be %d lower-case hex digits.", 2 * NCCL_ID_BYTES);
		return false;
	}
	return true;
}

/* called from _PG_init */
void
cbgpu_shim_define_gucs(void)
{
	DefineCustomStringVariable("cbgpu.nccl_id",
							   "Rendezvous token of the session's GPU interconnect (set by the dispatcher).",
							   NULL,
							   &cbgpu_nccl_id_hex,
							   "",
							   PGC_USERSET,
							   GUC_GPDB_NEED_SYNC | GUC_NO_SHOW_ALL | GUC_NOT_IN_SAMPLE,
							   check_nccl_id, NULL, NULL);
}

/*
 * QD side, before a plan with a replaced sub-tree is dispatched: make sure the session has a token and every QE knows it.
 */
void
cbgpu_shim_qd_prepare(bool previous_query_failed)
{
	unsigned char id[NCCL_ID_BYTES];
	char		hex[2 * NCCL_ID_BYTES + 1];
	StringInfoData cmd;

	if (Gp_role != GP_ROLE_DISPATCH)
		return;
	if (cbgpu_nccl_id_hex && cbgpu_nccl_id_hex[0] && !previous_query_failed)
		return;					/* the session's token stands */
	if (cbgpu_motion_unique_id(id) != CBGPU_OK)
		ereport(ERROR, (errcode(ERRCODE_INTERNAL_ERROR), errmsg("cbgpu: cannot create an interconnect token")));
	hex_encode((const char *) id, NCCL_ID_BYTES, hex);
	hex[2 * NCCL_ID_BYTES] = '\0';
	SetConfigOption("cbgpu.nccl_id", hex, PGC_USERSET, PGC_S_SESSION);
	initStringInfo(&cmd);
	appendStringInfo(&cmd, "SET cbgpu.nccl_id = '%s'", hex);
	CdbDispatchSetCommand(cmd.data, false);		/* QEs of gangs that already exist; later gangs get it from makeOptions */
	pfree(cmd.data);
}

/*
 * QE side: the interconnect for this query, or NULL on a single-segment cluster / the QD (a plan without Motions in the
 * replaced sub-tree needs none; one with Motions fails in cb_ExecInitNode with CBGPU_ERR_INVALID, and the shim leaves the
 * sub-tree to the CPU executor).
 */
CbInterconnect *
cbgpu_shim_interconnect(cbgpu_ctx *ctx, EState *estate)
{
	const int	nsegs = getgpsegmentCount();
	unsigned char id[NCCL_ID_BYTES];

	(void) estate;
	if (Gp_role != GP_ROLE_EXECUTE || nsegs < 2 || GpIdentity.segindex < 0)
		return NULL;
	if (cbgpu_nccl_id_hex == NULL || cbgpu_nccl_id_hex[0] == '\0')
		return NULL;
	if (session_ic != NULL && strcmp(current_id_hex, cbgpu_nccl_id_hex) == 0)
		return session_ic;
	if (session_ic != NULL)
	{
		/* the dispatcher replaced the token (it does after a query that died in error): the old communicator is abandoned on
		 * every segment alike, WITHOUT a collective step - a peer may have left an exchange half done, and
		 * cbgpu_motion_destroy's barrier would wait for it for ever (cbgpu_motion_abort: ncclCommAbort, unmap, free;
		 * the reference's TeardownInterconnect with hasErrors, cdb/ml_ipc.h:106) */
		cb_interconnect_destroy(session_ic);
		cbgpu_motion_abort(session_motion);
		session_ic = NULL;
		session_motion = NULL;
	}
	hex_decode(cbgpu_nccl_id_hex, 2 * NCCL_ID_BYTES, (char *) id);
	/* collective: every segment's QE is here for the same query.  Window size and the direct / staged choice are agreed
	 * inside; CHECK_FOR_INTERRUPTS cannot run inside NCCL's rendezvous, so cancel takes effect after it */
	if (cbgpu_motion_create(ctx, GpIdentity.segindex, nsegs, id, &session_motion) != CBGPU_OK)
		ereport(ERROR, (errcode(ERRCODE_GP_INTERCONNECTION_ERROR),
						errmsg("cbgpu: interconnect set-up failed: %s", cbgpu_last_error(ctx))));
	session_ic = cb_interconnect_nccl_create(session_motion);
	if (session_ic == NULL)
	{
		cbgpu_motion_destroy(session_motion);
		session_motion = NULL;
		ereport(ERROR, (errcode(ERRCODE_OUT_OF_MEMORY), errmsg("cbgpu: out of memory")));
	}
	strlcpy(current_id_hex, cbgpu_nccl_id_hex, sizeof(current_id_hex));
	return session_ic;
}
