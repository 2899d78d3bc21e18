package dgroomes.data_system_b200;

import dgroomes.data_system.Association;
import dgroomes.data_system.AssociationColumn;
import dgroomes.data_system.Column;
import dgroomes.data_system.Criteria;
import dgroomes.data_system.DataSystem;
import dgroomes.data_system.Query;
import dgroomes.data_system.QueryResult;
import dgroomes.data_system.Table;
import dgroomes.in_memory.InMemoryColumn;

import java.lang.foreign.Arena;
import java.lang.foreign.MemorySegment;
import java.nio.charset.StandardCharsets;
import java.util.ArrayDeque;
import java.util.BitSet;
import java.util.HashMap;
import java.util.IdentityHashMap;
import java.util.List;
import java.util.Map;
import java.util.Objects;

import static dgroomes.data_system_b200.ColqLibrary.*;
import static java.lang.foreign.ValueLayout.ADDRESS;
import static java.lang.foreign.ValueLayout.JAVA_BYTE;
import static java.lang.foreign.ValueLayout.JAVA_INT;
import static java.lang.foreign.ValueLayout.JAVA_LONG;

/**
 * {@code DataSystemSerialIndices} (data-system-serial-indices-arrays/.../DataSystemSerialIndices.java:14-102) on N GPUs
 * driven by ONE object in ONE JVM -- the shape of the reference, where {@code Runner} holds a single engine
 * (app/.../Runner.java:40).  One colq context per GPU, joined by {@code colq_comm_init_local} (peer access between the
 * GPUs, no process per GPU, no NCCL).
 * <p>
 * {@code register} takes the application's ordinary, whole tables.  At the first {@code execute} every table is split into
 * contiguous row ranges, one per GPU, with inner bounds at multiples of 64 rows ({@code colq_table_partition});
 * association columns keep the GLOBAL row indices the application wrote ({@code colq_associate_*_global}), so a hop may
 * leave the shard in either direction (bitmap all-gather / OR-reduce-scatter over NVLink inside libcolq).  The result is
 * the registered table's own {@code subset} of the matching global rows, exactly what the reference returns.
 * <p>
 * Python twin, exercised on 2 and 8 GPUs by tests/test_gpu_cross_shard.py: colq/local_group.py (DataSystemColqGroup).
 * NOTE: written against JDK 22 and never compiled in the build image (no JVM there).
 */
public final class DataSystemColqGroup implements DataSystem, AutoCloseable {

    private final int n;
    private final MemorySegment[] ctx;
    private final Arena arena = Arena.ofShared();
    private final MemorySegment ctxArray;
    private final Map<String, Table> tables = new HashMap<>();
    private final IdentityHashMap<Table, int[]> handles = new IdentityHashMap<>();      // per rank
    private final IdentityHashMap<Table, long[]> bounds = new IdentityHashMap<>();      // partition of a sharded table
    private final IdentityHashMap<Table, Integer> uploadedColumns = new IdentityHashMap<>();
    private final Map<String, Integer> registeredHandle = new HashMap<>();

    public DataSystemColqGroup(int... devices) {
        n = devices.length;
        ctx = new MemorySegment[n];
        ctxArray = arena.allocate(ADDRESS, n);
        try (Arena a = Arena.ofConfined()) {
            for (int r = 0; r < n; r++) {
                MemorySegment out = a.allocate(ADDRESS);
                int st = (int) colq_create.invokeExact(devices[r], out);
                if (st != OK) throw new IllegalStateException("colq_create(" + devices[r] + ") failed (" + st + "): libcolq has no CPU fallback");
                ctx[r] = out.get(ADDRESS, 0);
                ctxArray.setAtIndex(ADDRESS, r, ctx[r]);
            }
            check(0, (int) colq_comm_init_local.invokeExact(ctxArray, n));
        } catch (RuntimeException e) {
            throw e;
        } catch (Throwable t) {
            throw new IllegalStateException(t);
        }
    }

    /** DataSystemSerialIndices.register (:27-29): only records the reference; the split + upload happens at the first execute. */
    public synchronized void register(String tableName, Table table) {
        tables.put(tableName, table);
    }

    /** Contiguous, (almost) equal row ranges whose inner bounds are multiples of 64 rows (whole BitSet words). */
    static long[] evenPartition(long rows, int ranks) {
        long per = (rows + ranks - 1) / ranks;
        per = (per + 63) / 64 * 64;
        long[] b = new long[ranks + 1];
        for (int r = 0; r <= ranks; r++) b[r] = Math.min(r * per, rows);
        return b;
    }

    @Override
    public synchronized QueryResult execute(Query query) {
        Objects.requireNonNull(query, "The 'query' argument must not be null");
        if (!tables.containsKey(query.tableName))
            return new QueryResult.Failure("The query targets the table '%s' but that table is not registered".formatted(query.tableName));
        Table table = tables.get(query.tableName);
        try (Arena call = Arena.ofConfined()) {
            syncTables(call);
            MemorySegment qs = call.allocate(ADDRESS, n);
            try {
                for (int r = 0; r < n; r++) {
                    MemorySegment qOut = call.allocate(ADDRESS);
                    check(r, (int) colq_query_create.invokeExact(ctx[r], call.allocateFrom(query.tableName), qOut));
                    qs.setAtIndex(ADDRESS, r, qOut.get(ADDRESS, 0));
                    String opaque = translate(call, r, qOut.get(ADDRESS, 0), query);
                    if (opaque != null) return new QueryResult.Failure(opaque);
                }
                // every rank's work is enqueued before any rank's result is awaited (the kernels wait for one another ACROSS GPUs)
                int st = (int) colq_execute_group.invokeExact(ctxArray, qs, n);
                if (st == FAILURE) return new QueryResult.Failure(lastError(0));
                check(0, st);
                MemorySegment counts = call.allocate(JAVA_LONG, n);
                check(0, (int) colq_fetch_group.invokeExact(ctxArray, qs, n, counts));
                // a sharded root: every rank holds all ranks' GLOBAL rows, ascending; a replicated root: the same rows everywhere
                long count = counts.getAtIndex(JAVA_LONG, 0);
                MemorySegment idx = call.allocate(JAVA_INT, Math.max(count, 1));
                MemorySegment got = call.allocate(JAVA_LONG);
                check(0, (int) colq_fetch.invokeExact(ctx[0], qs.getAtIndex(ADDRESS, 0), MemorySegment.NULL, 0L, idx, count, got, MemorySegment.NULL));
                BitSet matchingRows = new BitSet(table.size());
                for (long i = 0; i < count; i++) matchingRows.set(idx.getAtIndex(JAVA_INT, i));
                return new QueryResult.Success(table.subset(matchingRows));   // DataSystemSerialIndices.java:100
            } finally {
                for (int r = 0; r < n; r++) {
                    MemorySegment q = qs.getAtIndex(ADDRESS, r);
                    if (!q.equals(MemorySegment.NULL)) { int ignored = (int) colq_query_destroy.invokeExact(q); }
                }
            }
        } catch (RuntimeException e) {
            throw e;
        } catch (Throwable t) {
            throw new IllegalStateException(t);
        }
    }

    /** Query tree -> colq_query of rank r; only structured predicates (Predicates.*): the tables are not dictionary-encoded here. */
    private String translate(Arena call, int r, MemorySegment q, Query query) throws Throwable {
        record Pending(Query.Node node, int id) {}
        var stack = new ArrayDeque<Pending>();
        stack.push(new Pending(query.rootNode, 0));
        while (!stack.isEmpty()) {
            Pending p = stack.pop();
            for (Criteria c : p.node().getCriteria()) {
                switch (c) {
                    case Criteria.IntCriteria(int ordinal, var pred) -> {
                        if (!(pred instanceof Predicates.IntRange(int lo, int hi)))
                            return "The criterion on ordinal %d is an opaque IntPredicate lambda; the GPU engine only runs structured predicates (dgroomes.data_system_b200.Predicates) and has no CPU fallback.".formatted(ordinal);
                        check(r, (int) colq_query_criteria_i32_range.invokeExact(q, p.id(), ordinal, lo, hi));
                    }
                    case Criteria.StringCriteria(int ordinal, var pred) -> {
                        if (!(pred instanceof Predicates.StringOp op))
                            return "The criterion on ordinal %d is an opaque Predicate<String> lambda; the GPU engine only runs structured predicates (dgroomes.data_system_b200.Predicates) and has no CPU fallback.".formatted(ordinal);
                        byte[] needle = op.needle();
                        MemorySegment nd = call.allocate(Math.max(needle.length, 1));
                        MemorySegment.copy(needle, 0, nd, JAVA_BYTE, 0, needle.length);
                        check(r, (int) colq_query_criteria_str.invokeExact(q, p.id(), ordinal, op.op().ordinal(), nd, needle.length));
                    }
                }
            }
            for (Map.Entry<Integer, Query.Node> e : p.node().getChildrenByOrdinal().entrySet()) {
                MemorySegment out = call.allocate(JAVA_INT);
                check(r, (int) colq_query_child.invokeExact(q, p.id(), (int) e.getKey(), out));
                stack.push(new Pending(e.getValue(), out.get(JAVA_INT, 0)));
            }
        }
        return null;
    }

    private void syncTables(Arena call) throws Throwable {
        var seen = new IdentityHashMap<Table, Boolean>();
        var todo = new ArrayDeque<>(tables.values());
        while (!todo.isEmpty()) {
            Table t = todo.pop();
            if (seen.put(t, true) != null) continue;
            for (Column c : t.columns()) if (c instanceof AssociationColumn ac) todo.push(ac.associatedEntity());
        }
        for (Table t : seen.keySet()) {
            if (handles.containsKey(t)) continue;
            long[] b = evenPartition(t.size(), n);
            MemorySegment bSeg = call.allocate(JAVA_LONG, n + 1L);
            for (int r = 0; r <= n; r++) bSeg.setAtIndex(JAVA_LONG, r, b[r]);
            int[] hs = new int[n];
            for (int r = 0; r < n; r++) {
                MemorySegment out = call.allocate(JAVA_INT);
                check(r, (int) colq_table_create.invokeExact(ctx[r], b[r + 1] - b[r], SHARDED, b[r], out));
                hs[r] = out.get(JAVA_INT, 0);
                check(r, (int) colq_table_partition.invokeExact(ctx[r], hs[r], bSeg, n));
            }
            handles.put(t, hs);
            bounds.put(t, b);
            uploadedColumns.put(t, 0);
        }
        for (Table t : seen.keySet()) {   // scalar columns: this rank's slice
            long[] b = bounds.get(t);
            List<? extends Column> cols = t.columns();
            for (int ordinal = uploadedColumns.get(t); ordinal < cols.size(); ordinal++) {
                for (int r = 0; r < n; r++) {
                    int h = handles.get(t)[r], lo = (int) b[r], cnt = (int) (b[r + 1] - b[r]);
                    switch (cols.get(ordinal)) {
                        case InMemoryColumn.IntegerColumn(int[] ints) -> {
                            MemorySegment seg = call.allocate(JAVA_INT, Math.max(cnt, 1));
                            MemorySegment.copy(ints, lo, seg, JAVA_INT, 0, cnt);
                            check(r, (int) colq_col_i32.invokeExact(ctx[r], h, ordinal, seg, (long) cnt));
                        }
                        case InMemoryColumn.StringColumn(String[] strings) -> {
                            byte[][] enc = new byte[cnt][];
                            long total = 0;
                            for (int i = 0; i < cnt; i++) { enc[i] = strings[lo + i].getBytes(StandardCharsets.UTF_8); total += enc[i].length; }
                            MemorySegment off = call.allocate(JAVA_INT, cnt + 1L), bytes = call.allocate(Math.max(total, 1));
                            long pos = 0;
                            for (int i = 0; i < cnt; i++) {
                                off.setAtIndex(JAVA_INT, i, (int) pos);
                                MemorySegment.copy(enc[i], 0, bytes, JAVA_BYTE, pos, enc[i].length);
                                pos += enc[i].length;
                            }
                            off.setAtIndex(JAVA_INT, cnt, (int) pos);
                            check(r, (int) colq_col_str.invokeExact(ctx[r], h, ordinal, off, bytes, (long) cnt, total));
                        }
                        case InMemoryColumn.BooleanColumn(boolean[] bools) -> {
                            MemorySegment seg = call.allocate(Math.max(cnt, 1));
                            for (int i = 0; i < cnt; i++) seg.set(JAVA_BYTE, i, (byte) (bools[lo + i] ? 1 : 0));
                            check(r, (int) colq_col_bool.invokeExact(ctx[r], h, ordinal, seg, (long) cnt));
                        }
                        default -> { /* association columns below */ }
                    }
                }
            }
        }
        for (Table t : seen.keySet()) {   // association pairs, uploaded once from the side that comes first
            long[] b = bounds.get(t);
            List<? extends Column> cols = t.columns();
            for (int ordinal = uploadedColumns.get(t); ordinal < cols.size(); ordinal++) {
                if (!(cols.get(ordinal) instanceof InMemoryColumn.AssociationColumn ac)) continue;
                Table y = ac.associatedEntity();
                int yOrdinal = identityIndexOf(y.columns(), ac.reverseAssociatedColumn());
                int h0 = handles.get(t)[0], hy0 = handles.get(y)[0];
                if (hy0 < h0 || (hy0 == h0 && yOrdinal < ordinal)) continue;
                Association[] assoc = ac.associations;
                for (int r = 0; r < n; r++) {
                    int lo = (int) b[r], cnt = (int) (b[r + 1] - b[r]);
                    long nnz = 0;
                    for (int i = 0; i < cnt; i++)
                        nnz += switch (assoc[lo + i]) { case Association.Many(int[] idx) -> idx.length; case Association.One o -> 1; case Association.None none -> 0; };
                    MemorySegment off = call.allocate(JAVA_LONG, cnt + 1L), tgt = call.allocate(JAVA_INT, Math.max(nnz, 1));
                    long pos = 0;
                    for (int i = 0; i < cnt; i++) {
                        off.setAtIndex(JAVA_LONG, i, pos);
                        switch (assoc[lo + i]) {
                            case Association.Many(int[] idx) -> { for (int v : idx) tgt.setAtIndex(JAVA_INT, pos++, v); }
                            case Association.One(int idx) -> tgt.setAtIndex(JAVA_INT, pos++, idx);
                            case Association.None ignored -> { }
                        }
                    }
                    off.setAtIndex(JAVA_LONG, cnt, pos);
                    // the keys the application wrote ARE global row indices of y
                    check(r, (int) colq_associate_csr_global.invokeExact(ctx[r], handles.get(t)[r], ordinal, handles.get(y)[r], yOrdinal, off, tgt, (long) cnt, nnz));
                }
            }
        }
        for (Table t : seen.keySet()) uploadedColumns.put(t, t.columns().size());
        for (var e : tables.entrySet()) {
            int[] hs = handles.get(e.getValue());
            if (!Integer.valueOf(hs[0]).equals(registeredHandle.get(e.getKey()))) {
                for (int r = 0; r < n; r++) check(r, (int) colq_register.invokeExact(ctx[r], call.allocateFrom(e.getKey()), hs[r]));
                registeredHandle.put(e.getKey(), hs[0]);
            }
        }
    }

    private static int identityIndexOf(List<? extends Column> cols, Object col) {
        for (int i = 0; i < cols.size(); i++) if (cols.get(i) == col) return i;
        throw new IllegalStateException("reverse association column is not a column of its table");
    }

    private String lastError(int r) throws Throwable {
        MemorySegment p = (MemorySegment) colq_last_error.invokeExact(ctx[r]);
        return p.reinterpret(1024).getString(0);
    }

    private void check(int r, int status) throws Throwable {
        switch (status) {
            case OK -> { }
            case THROW_INDEX_OOB -> throw new IndexOutOfBoundsException(lastError(r));
            case THROW_NULL -> throw new NullPointerException(lastError(r));
            case THROW_ILLEGAL_ARG -> throw new IllegalArgumentException(lastError(r));
            default -> throw new IllegalStateException("libcolq status " + status + " on rank " + r + ": " + lastError(r));
        }
    }

    @Override
    public synchronized void close() {
        try {
            for (int r = 0; r < n; r++) { int ignored = (int) colq_destroy.invokeExact(ctx[r]); }
        } catch (Throwable t) {
            throw new IllegalStateException(t);
        }
        arena.close();
    }
}
