// inserter_b200.go -- SOURCE ONLY: the cgo side of the drop-in, as a maintainer of
// cloudflare/flow-pipeline would add it next to inserter/inserter.go.  It cannot be
// compiled in this repository's build image (no Go toolchain, no sarama); the C++
// program in ../inserter.cc is the same loop, built and tested here, over the same
// C ABI (include/flowagg.h).
//
// What changes against inserter/inserter.go:
//   - state keeps one *C.fa_ctx per claimed partition instead of [][]interface{} and
//     the global mutex (inserter.go:75-88, :92, :115);
//   - buffer() memcpy's msg.Value into the library's pinned slab instead of
//     proto.Unmarshal + append (inserter.go:113-165);
//   - flush() emits flows_5m rows from fa_flush_begin / fa_flush_end instead of one INSERT per
//     flow (inserter.go:90-111), on -flush.dur and on -flush.count (inserter.go:36,118,161), without
//     stalling the claim while the table drains;
//   - MarkMessage moves after the slab hand-over (inserter.go:188).
// Flags, logging, metrics endpoint, consumer-group wiring in main() stay as they are.
//
//go:build cgo

package main

/*
#cgo CFLAGS: -I${SRCDIR}/../../../include
#cgo LDFLAGS: -L${SRCDIR}/.. -lflowagg
#include <stdlib.h>
#include <string.h>
#include "flowagg.h"
*/
import "C"

import (
	"flag"
	"fmt"
	"time"
	"unsafe"

	"github.com/Shopify/sarama"
	log "github.com/sirupsen/logrus"
)

// -offsets.gpu: hand fa_submit the length-delimited slab WITHOUT the offsets array; the library finds the record boundaries on
// the GPU (csrc/frame.cuh).  The slab crosses PCIe 4 bytes per flow lighter, which is throughput when the link is the limit
// (bench.py e2e: 0.639 against 0.619 G flows/s on one B200).  Safe when the shim writes the length prefixes itself (the bare
// topic, mocker.go:96-97); on the -proto.fixedlen topic a value whose own prefix lies would desynchronise the rest of its slab
// (every span is still validated: the damage is bad records, never a wrong row), so there the offsets stay the default.
var OffsetsOnGPU = flag.Bool("offsets.gpu", false, "find record boundaries on the GPU instead of shipping offsets")

type partitionState struct {
	ctx     *C.fa_ctx
	slot    C.int
	slab    *C.uint8_t
	offs    *C.uint32_t
	slabCap C.size_t
	recCap  C.size_t
	fill    C.size_t
	nrec    C.size_t
	pending []*sarama.ConsumerMessage

	sinceFlush int  // messages buffered since the last flush (-flush.count, inserter.go:118,161)
	draining   bool // fa_flush_begin issued, its fa_flush_end not yet
}

func newPartitionState(device int, keyMode C.uint32_t) *partitionState {
	var cfg C.fa_config
	cfg.abi_version = C.FA_ABI_VERSION
	cfg.device = C.int32_t(device)
	cfg.key_mode = keyMode
	ps := &partitionState{}
	if rc := C.fa_create(&cfg, &ps.ctx); rc != C.FA_OK {
		// no CPU fallback: the stage needs its GPU (reference style: log.Fatal, inserter.go:249)
		log.Fatalf("fa_create: %s (%s)", C.GoString(C.fa_strerror(rc)), C.GoString(C.fa_last_error(ps.ctx)))
	}
	ps.acquire()
	return ps
}

func (ps *partitionState) acquire() {
	if rc := C.fa_host_buffer(ps.ctx, ps.slot, &ps.slab, &ps.slabCap, &ps.offs, &ps.recCap); rc != C.FA_OK {
		log.Fatalf("fa_host_buffer: %s", C.GoString(C.fa_strerror(rc)))
	}
	ps.fill, ps.nrec = 0, 0
	*ps.offs = 0
}

func (ps *partitionState) submit(session sarama.ConsumerGroupSession) {
	if ps.nrec == 0 {
		return
	}
	offs := ps.offs
	if *OffsetsOnGPU {
		offs = nil // runs one call behind (include/flowagg.h); fa_flush*/fa_stats_get finish the last slab
	}
	if rc := C.fa_submit(ps.ctx, ps.slab, ps.fill, offs, C.uint32_t(ps.nrec), C.FA_FRAMED); rc != C.FA_OK {
		log.Fatalf("fa_submit: %s (%s)", C.GoString(C.fa_strerror(rc)), C.GoString(C.fa_last_error(ps.ctx)))
	}
	for _, m := range ps.pending { // offsets are marked once their bytes are in the stage
		session.MarkMessage(m, "")
	}
	Inserts.Add(float64(ps.nrec)) // the counter the reference registers but never increments (inserter.go:44-49)
	ps.pending = ps.pending[:0]
	ps.slot ^= 1
	ps.acquire() // blocks until that slab's previous host-to-device copy has finished
}

// buffer replaces (*state).buffer (inserter.go:113-165): the decode happens on the GPU.
func (ps *partitionState) buffer(session sarama.ConsumerGroupSession, msg *sarama.ConsumerMessage, fixedLen bool) {
	need := C.size_t(len(msg.Value) + 10)
	if ps.fill+need > ps.slabCap || ps.nrec >= ps.recCap {
		ps.submit(session)
	}
	dst := unsafe.Add(unsafe.Pointer(ps.slab), ps.fill)
	n := 0
	if !fixedLen { // bare value (mocker.go:96-97): add the varint length prefix ourselves
		l := uint64(len(msg.Value))
		b := (*[10]byte)(dst)
		for l >= 0x80 {
			b[n] = byte(l) | 0x80
			l >>= 7
			n++
		}
		b[n] = byte(l)
		n++
	}
	if len(msg.Value) > 0 { // sarama owns msg.Value only until the next receive: copy now
		C.memcpy(unsafe.Add(dst, n), unsafe.Pointer(&msg.Value[0]), C.size_t(len(msg.Value)))
	}
	ps.fill += C.size_t(n + len(msg.Value))
	ps.nrec++
	ps.sinceFlush++
	*(*C.uint32_t)(unsafe.Add(unsafe.Pointer(ps.offs), uintptr(ps.nrec)*4)) = C.uint32_t(ps.fill)
	ps.pending = append(ps.pending, msg)
}

// flush replaces (*state).flush (inserter.go:90-111): aggregate rows out, table reset.
func (ps *partitionState) flush(session sarama.ConsumerGroupSession, sink func(rows []C.fa_row)) {
	ps.submit(session)
	rows := make([]C.fa_row, 1<<16)
	var n C.size_t
	rc := C.fa_flush(ps.ctx, &rows[0], C.size_t(len(rows)), &n, 0)
	if rc == C.FA_ERR_CAPACITY {
		rows = make([]C.fa_row, n)
		rc = C.fa_flush(ps.ctx, &rows[0], n, &n, 0)
	}
	if rc != C.FA_OK && rc != C.FA_ERR_TABLE_FULL {
		log.Fatalf("fa_flush: %s (%s)", C.GoString(C.fa_strerror(rc)), C.GoString(C.fa_last_error(ps.ctx)))
	}
	sink(rows[:n])
}

// flushBegin / flushEnd: the same flush in two halves (include/flowagg.h).  Between them ConsumeClaim keeps buffering -- the
// GPU aggregates the new messages into a spare table while the filled one drains on a side stream -- where the
// reference's flush holds s.lock and stalls every partition (inserter.go:90-111).
func (ps *partitionState) flushBegin(session sarama.ConsumerGroupSession) {
	ps.submit(session)
	if rc := C.fa_flush_begin(ps.ctx, 0); rc != C.FA_OK {
		log.Fatalf("fa_flush_begin: %s (%s)", C.GoString(C.fa_strerror(rc)), C.GoString(C.fa_last_error(ps.ctx)))
	}
	ps.draining = true
	ps.sinceFlush = 0
}

func (ps *partitionState) flushEnd(sink func(rows []C.fa_row)) {
	if !ps.draining {
		return
	}
	ps.draining = false
	rows := make([]C.fa_row, 1<<16)
	var n C.size_t
	rc := C.fa_flush_end(ps.ctx, &rows[0], C.size_t(len(rows)), &n)
	if rc == C.FA_ERR_CAPACITY { // the rows are kept: once more with the size it reported
		rows = make([]C.fa_row, n)
		rc = C.fa_flush_end(ps.ctx, &rows[0], n, &n)
	}
	if rc != C.FA_OK && rc != C.FA_ERR_TABLE_FULL {
		log.Fatalf("fa_flush_end: %s (%s)", C.GoString(C.fa_strerror(rc)), C.GoString(C.fa_last_error(ps.ctx)))
	}
	sink(rows[:n])
}

// flushBox is the box-wide variant: Kafka does not partition by group key, so the partitions' tables hold
// partial sums of the same keys.  fa_flush_box exchanges them by key owner between the GPUs and returns
// each key once, in ORDER BY order (what the SummingMergeTree would converge to, create.sh:88-90).
// Call it with every live partitionState's ctx from one goroutine while the claims are quiescent
// (e.g. from Cleanup, inserter.go:172).
func flushBox(ctxs []*C.fa_ctx, sink func(rows []C.fa_row)) {
	rows := make([]C.fa_row, 1<<16)
	var n C.size_t
	rc := C.fa_flush_box(&ctxs[0], C.int(len(ctxs)), &rows[0], C.size_t(len(rows)), &n, 0)
	if rc == C.FA_ERR_CAPACITY {
		rows = make([]C.fa_row, n)
		rc = C.fa_flush_box(&ctxs[0], C.int(len(ctxs)), &rows[0], n, &n, 0)
	}
	if rc != C.FA_OK && rc != C.FA_ERR_TABLE_FULL {
		log.Fatalf("fa_flush_box: %s (%s)", C.GoString(C.fa_strerror(rc)), C.GoString(C.fa_last_error(ctxs[0])))
	}
	sink(rows[:n])
}

// ConsumeClaim replaces (*state).ConsumeClaim (inserter.go:176-196); sarama calls it once per
// claimed partition, each with its own fa_ctx: no cross-goroutine lock.
func (s *state) ConsumeClaimB200(session sarama.ConsumerGroupSession, claim sarama.ConsumerGroupClaim) error {
	ps := newPartitionState(int(claim.Partition())%s.gpus, C.FA_KEY_FLOWS5M)
	defer C.fa_destroy(ps.ctx)
	timer := time.After(*FlushTime)
	for {
		select {
		case message, ok := <-claim.Messages():
			if !ok {
				ps.flushEnd(s.writeRows) // a drain still in flight, then the last window in one go
				ps.flush(session, s.writeRows)
				return nil
			}
			log.Debugf("%s/%d/%d\t%s\t", message.Topic, message.Partition, message.Offset, message.Key)
			ps.buffer(session, message, s.fixedLen)
			// -flush.count (inserter.go:36,118,161): a flush every N buffered messages, timer or not
			if *FlushCount > 0 && ps.sinceFlush >= *FlushCount {
				ps.flushEnd(s.writeRows)
				ps.flushBegin(session)
			}
		case <-timer:
			ps.flushEnd(s.writeRows) // the previous window's rows: drained long ago
			ps.flushBegin(session)   // this window: swap tables and return to the claim at once
			timer = time.After(*FlushTime)
		}
	}
}

// writeRows is the sink: flows_5m rows (create.sh:70-87) as TSV; a Clickhouse RowBinary or
// Postgres COPY writer plugs in here.
func (s *state) writeRows(rows []C.fa_row) {
	for i := range rows {
		r := &rows[i]
		t := time.Unix(int64(r.key[0]), 0).UTC()
		fmt.Fprintf(s.out, "%s\t%s\t%d\t%d\t[%d]\t[%d]\t[%d]\t[%d]\t%d\t%d\t%d\n", t.Format("2006-01-02"),
			t.Format("2006-01-02 15:04:05"), r.key[1], r.key[2], r.key[3], r.bytes, r.packets, r.count, r.bytes, r.packets, r.count)
	}
}
