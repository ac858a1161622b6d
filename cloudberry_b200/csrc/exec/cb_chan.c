/*
 * cb_chan.c - packet channels between the processes of one query (include/cb_chan.h): the ring protocol, and the
 * POSIX shared-memory arenas.  The GPU peer-memory arenas live in csrc/motion.cu (cbgpu_motion_chan_mem).
 *
 * Stands where the reference's interconnects keep their per-connection packet queues and acks
 * (contrib/interconnect/udp/ic_udpifc.c: MotionConn ring buffers, handleAcks / sendAck; tcp/ic_tcp.c: readPacket).
 */
#define _GNU_SOURCE
#include <errno.h>
#include <fcntl.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <sys/mman.h>
#include <sys/stat.h>
#include <time.h>
#include <unistd.h>

#include "../../../include/cb_chan.h"

#define CH_ALIGN 256

struct cb_chan
{
	int			rank,
				nranks,
				slots,
				slot_bytes;		/* 8-byte header (length) + payload, a multiple of CH_ALIGN           */
	CbChanMem	mem;
	uint64_t   *sent;			/* [nranks] packets I completed into each destination's ring          */
	uint64_t   *acked;			/* [nranks] ... of which that destination is known to have consumed   */
	uint64_t   *consumed;		/* [nranks] packets I took from each sender's ring                    */
	uint64_t   *seen_tail;		/* [nranks] last tail[s] read from my arena                           */
	int			rr;				/* any-source receives start after the last source served             */
	unsigned char *stage;		/* one slot                                                           */
};

static size_t
ch_pad(size_t n)
{
	return (n + CH_ALIGN - 1) / CH_ALIGN * CH_ALIGN;
}

/* arena: [tail[nranks]] [ack[nranks]] (each padded) then ring[s] for every sender s */
static size_t
off_tail(const cb_chan *c, int s)
{
	(void) c;
	return (size_t) s * 8;
}

static size_t
off_ack(const cb_chan *c, int d)
{
	return ch_pad((size_t) c->nranks * 8) + (size_t) d * 8;
}

static size_t
off_slot(const cb_chan *c, int s, uint64_t seq)
{
	return 2 * ch_pad((size_t) c->nranks * 8) + ((size_t) s * (size_t) c->slots + (size_t) (seq % (uint64_t) c->slots)) * (size_t) c->slot_bytes;
}

size_t
cb_chan_arena_bytes(int nranks, int slots, int slot_bytes)
{
	return 2 * ch_pad((size_t) nranks * 8) + (size_t) nranks * (size_t) slots * ch_pad((size_t) slot_bytes);
}

cb_chan *
cb_chan_create(int rank, int nranks, int slots, int slot_bytes, const CbChanMem *mem)
{
	cb_chan    *c;

	if (nranks < 1 || rank < 0 || rank >= nranks || slots < 1 || slot_bytes < 64 || !mem || !mem->put || !mem->get)
		return NULL;
	c = calloc(1, sizeof(cb_chan));
	if (!c)
		return NULL;
	c->rank = rank;
	c->nranks = nranks;
	c->slots = slots;
	c->slot_bytes = (int) ch_pad((size_t) slot_bytes);
	c->mem = *mem;
	c->sent = calloc((size_t) nranks, 8);
	c->acked = calloc((size_t) nranks, 8);
	c->consumed = calloc((size_t) nranks, 8);
	c->seen_tail = calloc((size_t) nranks, 8);
	c->stage = malloc((size_t) c->slot_bytes);
	if (!c->sent || !c->acked || !c->consumed || !c->seen_tail || !c->stage)
	{
		cb_chan_destroy(c);
		return NULL;
	}
	return c;
}

void
cb_chan_destroy(cb_chan *c)
{
	if (!c)
		return;
	free(c->sent);
	free(c->acked);
	free(c->consumed);
	free(c->seen_tail);
	free(c->stage);
	free(c);
}

int
cb_chan_max_packet(const cb_chan *c)
{
	return c->slot_bytes - 8;
}

static int64_t
now_ms(void)
{
	struct timespec ts;

	clock_gettime(CLOCK_MONOTONIC, &ts);
	return (int64_t) ts.tv_sec * 1000 + ts.tv_nsec / 1000000;
}

static void
backoff(int spins)
{
	if (spins < 64)
		return;
	{
		struct timespec ts = {0, spins < 1024 ? 2000 : 50000};

		nanosleep(&ts, NULL);
	}
}

int
cb_chan_send(cb_chan *c, int dest, const void *pkt, int len, int timeout_ms)
{
	const int64_t t0 = now_ms();
	uint64_t	hdr;
	int			spins = 0;

	if (dest < 0 || dest >= c->nranks || len < 1 || len > c->slot_bytes - 8)
		return -1;
	/* flow control: the destination's ring of my packets must have a free slot */
	while (c->sent[dest] - c->acked[dest] >= (uint64_t) c->slots)
	{
		uint64_t	a = 0;

		if (c->mem.get(c->mem.arg, off_ack(c, dest), &a, 8) != 0)
			return -1;
		if (a > c->acked[dest])
		{
			c->acked[dest] = a;
			continue;
		}
		if (timeout_ms >= 0 && now_ms() - t0 >= timeout_ms)
			return 1;
		backoff(++spins);
	}
	hdr = (uint64_t) (uint32_t) len;
	memcpy(c->stage, &hdr, 8);
	memcpy(c->stage + 8, pkt, (size_t) len);
	if (c->mem.put(c->mem.arg, dest, off_slot(c, c->rank, c->sent[dest]), c->stage, ch_pad((size_t) len + 8)) != 0)
		return -1;
	c->sent[dest]++;
	/* the packet first, then the word that publishes it */
	if (c->mem.put(c->mem.arg, dest, off_tail(c, c->rank), &c->sent[dest], 8) != 0)
		return -1;
	return 0;
}

/* is a packet from s waiting?  (refreshes seen_tail[s] from my arena when the cached value says no) */
static int
ready(cb_chan *c, int s)
{
	if (c->seen_tail[s] > c->consumed[s])
		return 1;
	if (c->mem.get(c->mem.arg, off_tail(c, s), &c->seen_tail[s], 8) != 0)
		return -1;
	return c->seen_tail[s] > c->consumed[s];
}

int
cb_chan_pending(cb_chan *c, int src)
{
	int			n = 0;

	for (int s = 0; s < c->nranks; s++)
		if (src < 0 || s == src)
		{
			if (ready(c, s) < 0)
				return -1;
			n += (int) (c->seen_tail[s] - c->consumed[s]);
		}
	return n;
}

int
cb_chan_recv(cb_chan *c, int src, void *buf, int cap, int *from, int timeout_ms)
{
	const int64_t t0 = now_ms();
	int			spins = 0;

	if (src >= c->nranks)
		return -1;
	for (;;)
	{
		for (int k = 0; k < c->nranks; k++)
		{
			const int	s = src >= 0 ? src : (c->rr + 1 + k) % c->nranks;
			uint64_t	hdr = 0;
			int			r;
			int			len;

			if (src >= 0 && k > 0)
				break;
			r = ready(c, s);
			if (r < 0)
				return -1;
			if (!r)
				continue;
			if (c->mem.get(c->mem.arg, off_slot(c, s, c->consumed[s]), &hdr, 8) != 0)
				return -1;
			len = (int) (uint32_t) hdr;
			if (len < 1 || len > c->slot_bytes - 8)
				return -1;
			if (len > cap)
				return -2;
			if (c->mem.get(c->mem.arg, off_slot(c, s, c->consumed[s]) + 8, buf, (size_t) len) != 0)
				return -1;
			c->consumed[s]++;
			/* the slot is free again: tell the sender (its arena, my ack word) */
			if (c->mem.put(c->mem.arg, s, off_ack(c, c->rank), &c->consumed[s], 8) != 0)
				return -1;
			c->rr = s;
			if (from)
				*from = s;
			return len;
		}
		if (timeout_ms >= 0 && now_ms() - t0 >= timeout_ms)
			return 0;
		backoff(++spins);
	}
}

/* ------------------------------------------------------------------------------------------
 * POSIX shared-memory arenas
 * ------------------------------------------------------------------------------------------ */
#define SHM_HDR 64				/* [0]: 1 once the owner has mapped every peer's arena (cb_chan_shm_attach)              */

struct cb_chan_shm
{
	char		token[96];
	int			rank,
				nranks;
	size_t		bytes;
	unsigned char **map;		/* [nranks]: map[rank] is mine                                        */
};

static void
shm_name(const cb_chan_shm *s, int rank, char *out, size_t n)
{
	snprintf(out, n, "/%s.%d", s->token, rank);
}

cb_chan_shm *
cb_chan_shm_create(const char *token, int rank, int nranks, size_t arena_bytes)
{
	cb_chan_shm *s = calloc(1, sizeof(cb_chan_shm));
	char		name[128];
	int			fd;

	if (!s)
		return NULL;
	snprintf(s->token, sizeof(s->token), "%s", token);
	s->rank = rank;
	s->nranks = nranks;
	s->bytes = arena_bytes + SHM_HDR;
	s->map = calloc((size_t) nranks, sizeof(unsigned char *));
	shm_name(s, rank, name, sizeof(name));
	shm_unlink(name);
	fd = shm_open(name, O_CREAT | O_EXCL | O_RDWR, 0600);
	if (fd < 0 || ftruncate(fd, (off_t) s->bytes) != 0)
	{
		if (fd >= 0)
			close(fd);
		free(s->map);
		free(s);
		return NULL;
	}
	s->map[rank] = mmap(NULL, s->bytes, PROT_READ | PROT_WRITE, MAP_SHARED, fd, 0);
	close(fd);
	if (s->map[rank] == MAP_FAILED)
	{
		free(s->map);
		free(s);
		return NULL;
	}
	return s;					/* a fresh segment reads as zeros */
}

static int
shm_put(void *arg, int peer, size_t off, const void *src, size_t len)
{
	cb_chan_shm *s = arg;

	if (len == 8 && (off & 7) == 0)
	{
		/* a counter: published with release order, after the packet bytes stored before it */
		uint64_t	v;

		memcpy(&v, src, 8);
		__atomic_store_n((uint64_t *) (s->map[peer] + SHM_HDR + off), v, __ATOMIC_RELEASE);
		return 0;
	}
	memcpy(s->map[peer] + SHM_HDR + off, src, len);
	return 0;
}

static int
shm_get(void *arg, size_t off, void *dst, size_t len)
{
	cb_chan_shm *s = arg;

	if (len == 8 && (off & 7) == 0)
	{
		uint64_t	v = __atomic_load_n((uint64_t *) (s->map[s->rank] + SHM_HDR + off), __ATOMIC_ACQUIRE);

		memcpy(dst, &v, 8);
		return 0;
	}
	memcpy(dst, s->map[s->rank] + SHM_HDR + off, len);
	return 0;
}

/* 0 once every peer's arena is mapped here AND every peer has mapped everybody's (so that a fast endpoint, done and gone
 * and its name unlinked, can never be missed by a slow starter: nobody gets past this before everybody holds every mapping);
 * -1: not yet, call again */
int
cb_chan_shm_attach(cb_chan_shm *s, CbChanMem *mem)
{
	for (int p = 0; p < s->nranks; p++)
	{
		char		name[128];
		struct stat st;
		int			fd;

		if (p == s->rank || s->map[p])
			continue;
		shm_name(s, p, name, sizeof(name));
		fd = shm_open(name, O_RDWR, 0600);
		if (fd < 0)
			return -1;
		if (fstat(fd, &st) != 0 || (size_t) st.st_size < s->bytes)
		{
			close(fd);			/* created, not yet sized */
			return -1;
		}
		s->map[p] = mmap(NULL, s->bytes, PROT_READ | PROT_WRITE, MAP_SHARED, fd, 0);
		close(fd);
		if (s->map[p] == MAP_FAILED)
		{
			s->map[p] = NULL;
			return -1;
		}
	}
	__atomic_store_n((uint64_t *) s->map[s->rank], 1, __ATOMIC_RELEASE);
	for (int p = 0; p < s->nranks; p++)
		if (__atomic_load_n((uint64_t *) s->map[p], __ATOMIC_ACQUIRE) != 1)
			return -1;
	mem->arg = s;
	mem->put = shm_put;
	mem->get = shm_get;
	return 0;
}

void
cb_chan_shm_close(cb_chan_shm *s, int unlink_own)
{
	if (!s)
		return;
	for (int p = 0; p < s->nranks; p++)
		if (s->map[p])
			munmap(s->map[p], s->bytes);
	if (unlink_own)
	{
		char		name[128];

		shm_name(s, s->rank, name, sizeof(name));
		shm_unlink(name);
	}
	free(s->map);
	free(s);
}
