/*
 * cb_chan.h - packet channels between the processes of one query: the transport under the MotionIPCLayer
 * implementation integration/cbgpu_ic_layer.c.
 *
 * The reference's interconnects move "packets" of tuple chunks between QE processes over UDP / TCP sockets with their
 * own acks and flow control (contrib/interconnect/udp/ic_udpifc.c: ring of receive buffers per connection, capacity
 * Gp_interconnect_queue_depth; tcp/ic_tcp.c).  Here every process owns an ARENA that its peers can write into - a region
 * of its GPU's peer-memory window (stores travel over NVLink; cbgpu_motion_chan_mem, include/cbgpu.h) or, for hosts
 * without a GPU interconnect and for the CPU tests, a POSIX shared-memory segment (cb_chan_shm_*) - holding, per
 * sender, a ring of `slots` packets and two counters:
 *
 *     tail[s]   packets sender s has completed into ring[s]            (written by s, after the packet: release order)
 *     ack[d]    packets of MINE that destination d has consumed        (written by d into MY arena)
 *
 * A send waits while sent[d] - ack[d] == slots (the receiver's queue is full: the reference's flow control), puts the
 * packet into slot sent[d] % slots of ring[me] in d's arena, then publishes tail[me] = sent[d] + 1 there.  A receive
 * polls tail[] in its own arena, copies the oldest unread packet out and publishes ack[me] into the sender's arena.
 * No locks, no shared counters written by two parties; every word has one writer.
 */
#ifndef CB_CHAN_H
#define CB_CHAN_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

/* how an endpoint touches the arenas (its own and its peers'); every function returns 0 on success */
typedef struct CbChanMem
{
	void	   *arg;
	/* copy len bytes to offset off of rank `peer`'s arena; puts to one peer take effect in call order */
	int			(*put) (void *arg, int peer, size_t off, const void *src, size_t len);
	/* copy len bytes from offset off of MY arena (sees every put that was published before it by a later word) */
	int			(*get) (void *arg, size_t off, void *dst, size_t len);
} CbChanMem;

typedef struct cb_chan cb_chan;

/* bytes of arena every rank must provide for these parameters */
size_t		cb_chan_arena_bytes(int nranks, int slots, int slot_bytes);
/* the arena must be zero-filled on every rank before any rank sends */
cb_chan    *cb_chan_create(int rank, int nranks, int slots, int slot_bytes, const CbChanMem *mem);
void		cb_chan_destroy(cb_chan *c);
int			cb_chan_max_packet(const cb_chan *c);
/* 0: sent; 1: the receiver's ring stayed full for timeout_ms (nothing was sent); < 0: memory access failed */
int			cb_chan_send(cb_chan *c, int dest, const void *pkt, int len, int timeout_ms);
/* from rank `src`, or from whoever has a packet when src < 0 (round robin, so no sender starves: RecvTupleChunkFromAny's
 * fairness, ml_ipc.h:170-178).  Returns the packet's length (> 0) with *from set, 0 when nothing arrived within
 * timeout_ms, < 0 on error (-2: the caller's buffer is too small) */
int			cb_chan_recv(cb_chan *c, int src, void *buf, int cap, int *from, int timeout_ms);
/* packets a receive from `src` (or anyone) would find right now, without waiting */
int			cb_chan_pending(cb_chan *c, int src);

/* POSIX shared-memory arenas: every rank creates its own ("/<token>.<rank>"), then attaches to its peers' (all ranks
 * must have created theirs: the caller's rendezvous, e.g. the dispatcher's connection set-up).  The CbChanMem it fills
 * is valid until cb_chan_shm_close. */
typedef struct cb_chan_shm cb_chan_shm;
cb_chan_shm *cb_chan_shm_create(const char *token, int rank, int nranks, size_t arena_bytes);
/* 0: every peer's arena is mapped here AND every peer has mapped everybody's (from then on an endpoint may finish and
 * unlink its name without a slow starter missing it); -1: not yet - call again (with the caller's own time limit) */
int			cb_chan_shm_attach(cb_chan_shm *s, CbChanMem *mem);
void		cb_chan_shm_close(cb_chan_shm *s, int unlink_own);

#ifdef __cplusplus
}
#endif
#endif							/* CB_CHAN_H */
