/*
 * cb_aocs_load.c - one AOCS segment file set (all projected columns of one segno) from disk into a device relation.
 *
 * The storage side of aocs_beginscan / open_next_scan_seg (access/aocs/aocsam.c): for every projected column the
 * reference opens <relfilenode>[.<pseudo segno>] with pseudo segno = (filenum - 1) * 128 + segno
 * (FormatAOSegmentFileName, access/appendonly/aomd.c:84-117; AOTupleId_MultiplierSegmentFileNum,
 * access/appendonlytid.h:143) and reads it up to the EOF recorded for that column in pg_aocsseg.vpinfo
 * (AOCSVPInfoEntry.eof, cdb/cdbaocsam.h) -- bytes past it belong to aborted or in-progress appends.  Here the bytes go
 * to cbgpu_aocs_decode_column_ex (block walk, CRC-32C, decompression, datum stream decode on the device) and the
 * segment file's pg_aovisimap rows to cbgpu_aocs_apply_visimap.
 */
#include "../../../include/cb_exec.h"

#include <errno.h>
#include <fcntl.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <sys/stat.h>
#include <unistd.h>

#define CB_AOCS_FILENUM_MULT 128	/* AOTupleId_MultiplierSegmentFileNum */
#define CB_AOCS_MAX_SEGNO 127		/* AOTupleId_MaxSegmentFileNum        */

int
cb_aocs_segfile_path(const char *basepath, int segno, int filenum, char *out, size_t outsz)
{
	int			pseudo;
	int			n;

	if (!basepath || !out || segno < 0 || segno > CB_AOCS_MAX_SEGNO || filenum < 1)
		return -1;
	pseudo = (filenum - 1) * CB_AOCS_FILENUM_MULT + segno;
	if (pseudo > 0)
		n = snprintf(out, outsz, "%s.%u", basepath, (unsigned) pseudo);
	else
		n = snprintf(out, outsz, "%s", basepath);
	return (n < 0 || (size_t) n >= outsz) ? -1 : 0;
}

static int
load_fail(char *err, size_t errsz, int code, const char *fmt, const char *path, long long v)
{
	if (err && errsz)
		snprintf(err, errsz, fmt, path, v);
	return code;
}

/* the first `eof` bytes of the file (eof < 0: all of it) */
static int
read_prefix(const char *path, int64_t eof, unsigned char **bufp, int64_t *lenp, char *err, size_t errsz)
{
	struct stat st;
	unsigned char *buf;
	int64_t		want,
				got = 0;
	int			fd = open(path, O_RDONLY);

	if (fd < 0)
		return load_fail(err, errsz, CBGPU_ERR_INVALID, "cannot open segment file %s (errno %lld)", path, (long long) errno);
	if (fstat(fd, &st) != 0)
	{
		close(fd);
		return load_fail(err, errsz, CBGPU_ERR_INVALID, "cannot stat segment file %s (errno %lld)", path, (long long) errno);
	}
	want = eof < 0 ? (int64_t) st.st_size : eof;
	if (want > (int64_t) st.st_size)
	{
		close(fd);
		return load_fail(err, errsz, CBGPU_ERR_INVALID, "segment file %s is shorter than its recorded EOF %lld", path, (long long) eof);
	}
	buf = malloc((size_t) want + 16);
	if (!buf)
	{
		close(fd);
		return CBGPU_ERR_NOMEM;
	}
	while (got < want)
	{
		const ssize_t r = pread(fd, buf + got, (size_t) (want - got), (off_t) got);

		if (r <= 0)
		{
			if (r < 0 && errno == EINTR)
				continue;
			free(buf);
			close(fd);
			return load_fail(err, errsz, CBGPU_ERR_INVALID, "short read from segment file %s at offset %lld", path, (long long) got);
		}
		got += r;
	}
	close(fd);
	*bufp = buf;
	*lenp = want;
	return CBGPU_OK;
}

int
cb_aocs_dict_collect_segfile(cbgpu_ctx *ctx, const char *basepath, int segno, int checksum, const CbAocsColumnSpec *spec, char *err,
							 size_t errsz)
{
	char		path[4096];
	unsigned char *buf = NULL;
	int64_t		len = 0;
	int			rc;

	if (err && errsz)
		err[0] = 0;
	if (!spec || spec->varkind != CBGPU_AOCS_VAR_DICT || !spec->dict ||
		cb_aocs_segfile_path(basepath, segno, spec->filenum, path, sizeof(path)) != 0)
		return load_fail(err, errsz, CBGPU_ERR_INVALID, "bad dictionary column specification for %s (%lld)", basepath ? basepath : "", (long long) segno);
	rc = read_prefix(path, spec->eof, &buf, &len, err, errsz);
	if (rc != CBGPU_OK)
		return rc;
	rc = cbgpu_aocs_dict_collect(ctx, buf, len, checksum, spec->compresstype, spec->typalign, spec->dict);
	if (rc != CBGPU_OK && err && errsz)
		snprintf(err, errsz, "%s: %s", path, cbgpu_last_error(ctx));
	free(buf);
	return rc;
}

int
cb_aocs_load_segfile(cbgpu_ctx *ctx, const char *basepath, int segno, int checksum, int ncols, const CbAocsColumnSpec *cols,
					 cbgpu_rel *rel, int64_t row_offset, const cbgpu_visimap_entry *entries, int nentries, int64_t *nrows_out,
					 int64_t *nhidden_out, char *err, size_t errsz)
{
	int64_t		nrows = -1;
	unsigned char *first = NULL;
	int64_t		firstlen = 0;
	int			rc = CBGPU_OK;

	if (nrows_out)
		*nrows_out = 0;
	if (nhidden_out)
		*nhidden_out = 0;
	if (err && errsz)
		err[0] = 0;
	if (ncols < 1 || !cols)
		return load_fail(err, errsz, CBGPU_ERR_INVALID, "no columns to load from %s (%lld)", basepath ? basepath : "", (long long) ncols);
	for (int c = 0; c < ncols && rc == CBGPU_OK; c++)
	{
		char		path[4096];
		unsigned char *buf = NULL;
		int64_t		len = 0,
					n = 0;

		if (cb_aocs_segfile_path(basepath, segno, cols[c].filenum, path, sizeof(path)) != 0)
		{
			rc = load_fail(err, errsz, CBGPU_ERR_INVALID, "bad segment file name for %s (column %lld)", basepath ? basepath : "", (long long) c);
			break;
		}
		rc = read_prefix(path, cols[c].eof, &buf, &len, err, errsz);
		if (rc != CBGPU_OK)
			break;
		if (cols[c].varkind == CBGPU_AOCS_VAR_DICT)
			rc = cbgpu_aocs_decode_dict_column(ctx, buf, len, checksum, cols[c].compresstype, cols[c].typalign, cols[c].dict, rel,
											   cols[c].relcol, row_offset, &n);
		else
			rc = cbgpu_aocs_decode_column_ex(ctx, buf, len, checksum, cols[c].compresstype, cols[c].attlen, cols[c].varkind,
											 cols[c].typalign, rel, cols[c].relcol, row_offset, &n);
		if (rc != CBGPU_OK)
		{
			if (err && errsz)
				snprintf(err, errsz, "%s: %s", path, cbgpu_last_error(ctx));
		}
		else if (nrows >= 0 && n != nrows)
			rc = load_fail(err, errsz, CBGPU_ERR_INVALID, "column file %s holds %lld rows, the segment file's other columns differ", path, (long long) n);
		else
			nrows = n;
		if (c == 0 && rc == CBGPU_OK)
		{
			first = buf;		/* kept for the visibility map: its block headers give the row numbers */
			firstlen = len;
		}
		else
			free(buf);
	}
	if (rc == CBGPU_OK && (nentries > 0 || cbgpu_rel_visimap_dev(rel)) && nrows > 0)
	{
		rc = cbgpu_aocs_apply_visimap(ctx, first, firstlen, checksum, entries, nentries, rel, row_offset, nhidden_out);
		if (rc != CBGPU_OK && err && errsz)
			snprintf(err, errsz, "%s visimap: %s", basepath, cbgpu_last_error(ctx));
	}
	free(first);
	if (rc == CBGPU_OK && nrows_out)
		*nrows_out = nrows < 0 ? 0 : nrows;
	return rc;
}
