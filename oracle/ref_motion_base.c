/*
 * ref_motion_base.c - palloc / ereport for oracle/_ref/libmotion_ref.so (test infrastructure): what ref_aocs.c provides for
 * libaocs_ref.so, without that file's AOCS writer.  ereport(ERROR) unwinds to the driver's setjmp with the message kept for
 * ref_aocs_last_error().  The virtual-slot pieces SerializeTuple needs are restated as in ref_tupser.c.
 */
#include "postgres.h"

#include <setjmp.h>

#include "access/htup_details.h"
#include "executor/tuptable.h"

/* port.h maps the printf family onto pg_*printf, which are what this file DEFINES over the C library's */
#undef snprintf
#undef vsnprintf
#undef sprintf
#undef vsprintf

jmp_buf		ref_jmp;
static char ref_errbuf[512];
static int	ref_elevel;
const TupleTableSlotOps TTSOpsVirtual;

void
ref_abort(const char *what)
{
	snprintf(ref_errbuf, sizeof(ref_errbuf), "oracle/_ref: %s is a stub", what);
	longjmp(ref_jmp, 1);
}

const char *
ref_aocs_last_error(void)
{
	return ref_errbuf;
}

void	   *palloc(Size size) { return malloc(size ? size : 1); }
void	   *palloc0(Size size) { return calloc(1, size ? size : 1); }
void	   *repalloc(void *p, Size size) { return realloc(p, size ? size : 1); }
void		pfree(void *p) { free(p); }

bool
errstart(int elevel, const char *domain)
{
	(void) domain;
	ref_elevel = elevel;
	return elevel >= ERROR;		/* LOG / DEBUG chatter is dropped */
}

bool
errstart_cold(int elevel, const char *domain)
{
	return errstart(elevel, domain);
}

void
errfinish(const char *filename, int lineno, const char *funcname)
{
	(void) funcname;
	if (ref_elevel >= ERROR)
	{
		size_t		n = strlen(ref_errbuf);

		snprintf(ref_errbuf + n, sizeof(ref_errbuf) - n, " (%s:%d)", filename, lineno);
		longjmp(ref_jmp, 1);
	}
}

void
errmsg(const char *fmt,...)
{
	va_list		ap;

	va_start(ap, fmt);
	vsnprintf(ref_errbuf, sizeof(ref_errbuf), fmt, ap);
	va_end(ap);
}

void
errmsg_internal(const char *fmt,...)
{
	va_list		ap;

	va_start(ap, fmt);
	vsnprintf(ref_errbuf, sizeof(ref_errbuf), fmt, ap);
	va_end(ap);
}

void		errdetail(const char *fmt,...) { (void) fmt; }
void		errdetail_internal(const char *fmt,...) { (void) fmt; }
void		errcode(int sqlerrcode) { (void) sqlerrcode; }
void		errcode_for_file_access(void) {}

int
pg_snprintf(char *str, size_t count, const char *fmt,...)
{
	va_list		ap;
	int			n;

	va_start(ap, fmt);
	n = vsnprintf(str, count, fmt, ap);
	va_end(ap);
	return n;
}

MinimalTuple
ExecFetchSlotMinimalTuple(TupleTableSlot *slot, bool *shouldFree)
{
	*shouldFree = true;
	return heap_form_minimal_tuple(slot->tts_tupleDescriptor, slot->tts_values, slot->tts_isnull);
}

int
pg_vsnprintf(char *str, size_t count, const char *fmt, va_list args)
{
	return vsnprintf(str, count, fmt, args);
}
