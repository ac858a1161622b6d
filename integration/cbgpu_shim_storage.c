/*
 * cbgpu_shim_storage.c - the storage half of the backend-side shim: cbgpu_shim_load_relation(), the hook
 * integration/cbgpu_shim.c calls for every AOCS relation under a replaced sub-tree.
 *
 * It does with catalog lookups what aocs_beginscan / open_next_scan_seg do before the first block is read
 * (access/aocs/aocsam.c:548-700): the segment files and per-column EOFs of the relation (GetAllAOCSFileSegInfo,
 * access/aocs/aocssegfiles.c), each column's file number (GetFilenumForAttribute, catalog/pg_attribute_encoding.c) and
 * storage options (RelationGetAttributeOptions), the base path (relpathbackend), the visibility map rows
 * (pg_aovisimap_<oid>, catalog/aovisimap.h) -- and hands them to cb_aocs_load_segfile (include/cb_exec.h), which reads the
 * files and decodes them on the device.  No datum is decoded by the backend.
 *
 * Like cbgpu_shim.c this file is type-checked against the reference's headers (tests/test_shim_compiles.py); it cannot
 * be linked or run in this repository's container (no backend).
 */
#include "postgres.h"

#include "access/aocssegfiles.h"
#include "access/aosegfiles.h"
#include "access/genam.h"
#include "access/htup_details.h"
#include "access/table.h"
#include "catalog/aovisimap.h"
#include "catalog/pg_appendonly.h"
#include "catalog/pg_attribute_encoding.h"
#include "catalog/pg_type.h"
#include "common/relpath.h"
#include "fmgr.h"
#include "nodes/pg_list.h"
#include "utils/memutils.h"
#include "utils/rel.h"
#include "utils/snapmgr.h"

#include "cb_exec.h"

/* dictionaries of the string columns loaded so far, for translating the plan's Const nodes (cbgpu_dict_lookup) and the
 * result codes back (cbgpu_dict_entry) */
typedef struct ShimDict
{
	Oid			relid;
	AttrNumber	attno;
	cbgpu_dict *dict;
} ShimDict;
static List *shim_dicts = NIL;

cbgpu_dict *
cbgpu_shim_dict(Oid relid, AttrNumber attno)
{
	ListCell   *lc;

	foreach(lc, shim_dicts)
	{
		ShimDict   *d = (ShimDict *) lfirst(lc);

		if (d->relid == relid && d->attno == attno)
			return d->dict;
	}
	return NULL;
}

/* pg_attribute row -> column type of the device relation + how its datums are stored */
static void
column_shape(Form_pg_attribute att, int32 *cbtype, int32 *dscale, CbAocsColumnSpec *spec)
{
	spec->attlen = att->attlen;
	spec->varkind = 0;
	spec->typalign = att->attalign == 'c' ? 1 : att->attalign == 's' ? 2 : att->attalign == 'i' ? 4 : 8;
	*dscale = 0;
	switch (att->atttypid)
	{
		case INT4OID: *cbtype = CB_INT4; break;
		case INT8OID: *cbtype = CB_INT8; break;
		case DATEOID: *cbtype = CB_DATE; break;
		case FLOAT8OID: *cbtype = CB_FLOAT8; break;
		case BOOLOID: *cbtype = CB_BOOL; break;
		case NUMERICOID:
			if (att->atttypmod < (int32) VARHDRSZ)
				ereport(ERROR, (errcode(ERRCODE_FEATURE_NOT_SUPPORTED), errmsg("cbgpu: numeric column without a declared scale")));
			*cbtype = CB_NUMERIC;
			*dscale = (att->atttypmod - VARHDRSZ) & 0xffff;		/* numeric typmod: precision << 16 | scale */
			spec->varkind = CBGPU_AOCS_VAR_NUMERIC;
			break;
		case BPCHAROID:
			if (att->atttypmod - (int32) VARHDRSZ == 1)
			{
				*cbtype = CB_BPCHAR1;
				spec->varkind = CBGPU_AOCS_VAR_BPCHAR1;
				break;
			}
			/* fall through */
		case VARCHAROID:
		case TEXTOID:
			*cbtype = CB_DICT32;
			spec->varkind = CBGPU_AOCS_VAR_DICT;
			break;
		default:
			ereport(ERROR, (errcode(ERRCODE_FEATURE_NOT_SUPPORTED),
							errmsg("cbgpu: column type %u is not handled by the device path", att->atttypid)));
	}
}

static int32
column_compression(const StdRdOptions *opt)
{
	if (opt->compresstype[0] == '\0' || pg_strcasecmp(opt->compresstype, "none") == 0)
		return CBGPU_AOCS_COMPRESS_NONE;
	if (pg_strcasecmp(opt->compresstype, "zlib") == 0)
		return CBGPU_AOCS_COMPRESS_ZLIB;
	if (pg_strcasecmp(opt->compresstype, "zstd") == 0)
		return CBGPU_AOCS_COMPRESS_ZSTD;
	if (pg_strcasecmp(opt->compresstype, "rle_type") == 0)	/* levels 2-4 add zlib (init_datumstream_info, datumstream.c:396-436) */
		return opt->compresslevel >= 2 ? CBGPU_AOCS_COMPRESS_ZLIB : CBGPU_AOCS_COMPRESS_NONE;
	ereport(ERROR, (errcode(ERRCODE_FEATURE_NOT_SUPPORTED), errmsg("cbgpu: compresstype \"%s\" is not decoded on the device", opt->compresstype)));
	return 0;
}

/* the pg_aovisimap rows of one segment file, as cbgpu_visimap_entry (payloads detoasted into the current context) */
static int
visimap_entries(Oid visimaprelid, Snapshot snapshot, int32 segno, cbgpu_visimap_entry **out)
{
	Relation	vm = table_open(visimaprelid, AccessShareLock);
	SysScanDesc scan = systable_beginscan(vm, InvalidOid, false, snapshot, 0, NULL);
	HeapTuple	tup;
	int			n = 0,
				cap = 16;
	cbgpu_visimap_entry *e = palloc(sizeof(cbgpu_visimap_entry) * cap);

	while ((tup = systable_getnext(scan)) != NULL)
	{
		bool		isnull;
		Datum		d = heap_getattr(tup, Anum_pg_aovisimap_segno, RelationGetDescr(vm), &isnull);

		if (isnull || DatumGetInt32(d) != segno)
			continue;
		if (n == cap)
			e = repalloc(e, sizeof(cbgpu_visimap_entry) * (cap *= 2));
		e[n].first_row_num = DatumGetInt64(heap_getattr(tup, Anum_pg_aovisimap_firstrownum, RelationGetDescr(vm), &isnull));
		d = heap_getattr(tup, Anum_pg_aovisimap_visimap, RelationGetDescr(vm), &isnull);
		if (isnull)
		{
			e[n].data = NULL;	/* all visible (AppendOnlyVisimapEntry_Copyout, appendonly_visimap_entry.c:222-229) */
			e[n].len = 0;
		}
		else
		{
			struct varlena *v = pg_detoast_datum_copy((struct varlena *) DatumGetPointer(d));

			e[n].data = VARDATA(v);				/* int32 version + Bitmap_Compress output */
			e[n].len = (int32) (VARSIZE(v) - VARHDRSZ);
		}
		n++;
	}
	systable_endscan(scan);
	table_close(vm, AccessShareLock);
	*out = e;
	return n;
}

cbgpu_rel *
cbgpu_shim_load_relation(cbgpu_ctx *ctx, Relation rel, List *projected_attnos)
{
	TupleDesc	td = RelationGetDescr(rel);
	const int	ncols = projected_attnos ? list_length(projected_attnos) : td->natts;
	Snapshot	snapshot = GetActiveSnapshot();
	StdRdOptions **opts = RelationGetAttributeOptions(rel);
	CbAocsColumnSpec *spec = palloc0(sizeof(CbAocsColumnSpec) * ncols);
	int32	   *types = palloc(sizeof(int32) * ncols);
	int32	   *dscales = palloc(sizeof(int32) * ncols);
	AttrNumber *attno = palloc(sizeof(AttrNumber) * ncols);
	AOCSFileSegInfo **segs;
	int			nsegs = 0;
	int64		total = 0,
				row_offset = 0;
	char	   *basepath = relpathbackend(rel->rd_node, rel->rd_backend, MAIN_FORKNUM);
	Oid			visimaprelid = InvalidOid;
	bool		checksum = false;
	cbgpu_rel  *out = NULL;
	char		err[512];
	ListCell   *lc;
	int			c = 0;

	if (projected_attnos)
		foreach(lc, projected_attnos)
			attno[c++] = (AttrNumber) lfirst_int(lc);
	else
		for (c = 0; c < ncols; c++)
			attno[c] = (AttrNumber) (c + 1);
	for (c = 0; c < ncols; c++)
	{
		Form_pg_attribute att = TupleDescAttr(td, attno[c] - 1);

		column_shape(att, &types[c], &dscales[c], &spec[c]);
		spec[c].relcol = c;
		spec[c].filenum = GetFilenumForAttribute(RelationGetRelid(rel), attno[c]);
		spec[c].compresstype = column_compression(opts[attno[c] - 1]);
		checksum = opts[attno[c] - 1]->checksum;	/* a table-level option: the same for every column */
	}
	GetAppendOnlyEntryAuxOids(rel, NULL, NULL, NULL, &visimaprelid, NULL);
	segs = GetAllAOCSFileSegInfo(rel, snapshot, &nsegs, NULL);
	for (int s = 0; s < nsegs; s++)
		if (segs[s]->state == AOSEG_STATE_DEFAULT)
			total += segs[s]->total_tupcount;
	if (cbgpu_rel_create(ctx, total, ncols, types, dscales, &out) != CBGPU_OK)
		ereport(ERROR, (errcode(ERRCODE_INTERNAL_ERROR), errmsg("cbgpu: %s", cbgpu_last_error(ctx))));

	/* string columns: first pass over every segment file for the distinct values */
	for (c = 0; c < ncols; c++)
	{
		ShimDict   *sd;
		MemoryContext old;

		if (spec[c].varkind != CBGPU_AOCS_VAR_DICT)
			continue;
		if (cbgpu_dict_create(ctx, 1 << 20, (int64) 256 << 20, TupleDescAttr(td, attno[c] - 1)->atttypid == BPCHAROID, &spec[c].dict) != CBGPU_OK)
		{
			cbgpu_rel_free(out);
			ereport(ERROR, (errcode(ERRCODE_INTERNAL_ERROR), errmsg("cbgpu: %s", cbgpu_last_error(ctx))));
		}
		for (int s = 0; s < nsegs; s++)
		{
			if (segs[s]->state != AOSEG_STATE_DEFAULT || segs[s]->total_tupcount == 0)
				continue;
			spec[c].eof = getAOCSVPEntry(segs[s], attno[c] - 1)->eof;
			if (cb_aocs_dict_collect_segfile(ctx, basepath, segs[s]->segno, checksum, &spec[c], err, sizeof(err)) != CBGPU_OK)
			{
				cbgpu_dict_free(spec[c].dict);
				cbgpu_rel_free(out);
				ereport(ERROR, (errcode(ERRCODE_DATA_CORRUPTED), errmsg("cbgpu: %s", err)));
			}
		}
		if (cbgpu_dict_finalize(spec[c].dict, NULL) != CBGPU_OK)
		{
			cbgpu_dict_free(spec[c].dict);
			cbgpu_rel_free(out);
			ereport(ERROR, (errcode(ERRCODE_INTERNAL_ERROR), errmsg("cbgpu: %s", cbgpu_last_error(ctx))));
		}
		/* a reload of the relation supersedes the older dictionary of this column: give its 256 MB arena back */
		{
			ListCell   *dl;

			foreach(dl, shim_dicts)
			{
				ShimDict   *od = (ShimDict *) lfirst(dl);

				if (od->relid == RelationGetRelid(rel) && od->attno == attno[c])
				{
					cbgpu_dict_free(od->dict);
					shim_dicts = foreach_delete_current(shim_dicts, dl);
					pfree(od);
				}
			}
		}
		old = MemoryContextSwitchTo(TopMemoryContext);
		sd = palloc(sizeof(ShimDict));
		sd->relid = RelationGetRelid(rel);
		sd->attno = attno[c];
		sd->dict = spec[c].dict;
		shim_dicts = lcons(sd, shim_dicts);	/* newest first: a reload of the relation supersedes the older dictionary */
		MemoryContextSwitchTo(old);
	}

	/* every live segment file: read, decode, hide */
	for (int s = 0; s < nsegs; s++)
	{
		cbgpu_visimap_entry *entries = NULL;
		int			nentries;
		int64		nrows = 0,
					nhidden = 0;

		if (segs[s]->state != AOSEG_STATE_DEFAULT || segs[s]->total_tupcount == 0)
			continue;
		for (c = 0; c < ncols; c++)
			spec[c].eof = getAOCSVPEntry(segs[s], attno[c] - 1)->eof;
		nentries = visimap_entries(visimaprelid, snapshot, segs[s]->segno, &entries);
		if (cb_aocs_load_segfile(ctx, basepath, segs[s]->segno, checksum, ncols, spec, out, row_offset, entries, nentries, &nrows, &nhidden,
								 err, sizeof(err)) != CBGPU_OK)
		{
			cbgpu_rel_free(out);
			ereport(ERROR, (errcode(ERRCODE_DATA_CORRUPTED), errmsg("cbgpu: %s", err)));
		}
		if (nrows != segs[s]->total_tupcount)
		{
			cbgpu_rel_free(out);
			ereport(ERROR, (errcode(ERRCODE_DATA_CORRUPTED),
							errmsg("cbgpu: segment file %d of \"%s\" holds " INT64_FORMAT " rows, pg_aocsseg says " INT64_FORMAT,
								   segs[s]->segno, RelationGetRelationName(rel), nrows, segs[s]->total_tupcount)));
		}
		row_offset += nrows;
	}
	cbgpu_rel_set_nrows(out, row_offset);
	FreeAllAOCSSegFileInfo(segs, nsegs);
	pfree(basepath);
	return out;
}
