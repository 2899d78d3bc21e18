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
import java.nio.ByteOrder;
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
 * The B200 execution module behind the reference's {@link DataSystem} interface: same two methods as
 * {@code DataSystemSerialIndices} (data-system-serial-indices-arrays/.../DataSystemSerialIndices.java:27,53).
 * <p>
 * Host work is limited to (1) flattening the registered {@link Table} graph into off-heap {@link MemorySegment}
 * column buffers and handing them to libcolq.so once, (2) translating the {@link Query} tree into
 * {@code colq_query_*} downcalls, (3) turning the returned BitSet words into the result table with the registered
 * table's own {@code subset}. Everything else happens on the GPU. There is no CPU fallback.
 * <p>
 * NOTE: written against JDK 22 and never compiled in the build image (no JVM there); the ctypes twin of this class,
 * colq/engine.py, is the one exercised by the test suite through the very same C ABI.
 */
public final class DataSystemColq implements DataSystem, AutoCloseable {

    private final Map<String, Table> tables = new HashMap<>();
    private final IdentityHashMap<Table, Integer> handles = new IdentityHashMap<>();
    private final IdentityHashMap<Table, Integer> uploadedColumns = new IdentityHashMap<>();
    private final Map<String, Integer> registeredHandle = new HashMap<>();
    private final Arena arena = Arena.ofShared();
    private final MemorySegment ctx;

    /**
     * Physical layout on the engine side; results are identical in every combination.
     * hostResident: int / string / to-one association columns live in pinned off-heap segments (colq_host_alloc) that the
     * kernels read in place over PCIe -- nothing is copied at registration, a query moves only what it touches, and the
     * first full scan of a column leaves a copy in HBM.  dictionary: string columns are stored as int32 codes + distinct
     * values, and ANY Predicate&lt;String&gt; -- including the app's unchanged lambdas (Runner.java:236,255-259) -- is
     * evaluated once per distinct value here and applied on the GPU as a code lookup.
     */
    public record Layout(boolean hostResident, boolean dictionary, boolean deviceIngest) {
        public Layout(boolean hostResident, boolean dictionary) { this(hostResident, dictionary, false); }
        /** Everything the unchanged app needs: pinned off-heap columns, string AND integer columns dictionary-encoded. */
        public static Layout forUnchangedLambdas() { return new Layout(true, true, false); }
        /**
         * The same, with the load-time work done by the GPU (include/colq.h "Ingest on the device"): string columns are
         * shipped as plain offsets + bytes and dictionary-encoded by colq_col_str_encode, every association goes up as one
         * CSR that colq_associate validates and classifies into dense to-one vs to-many.
         */
        public static Layout deviceIngest() { return new Layout(false, true, true); }
    }

    private final Layout layout;
    private final java.util.ArrayList<MemorySegment> hostBuffers = new java.util.ArrayList<>();
    private final IdentityHashMap<Table, Map<Integer, String[]>> dictionaryValues = new IdentityHashMap<>();
    private final IdentityHashMap<Table, Map<Integer, int[]>> intDictionaryValues = new IdentityHashMap<>();

    public DataSystemColq() {
        this(0, new Layout(false, false));
    }

    public DataSystemColq(int device) {
        this(device, new Layout(false, false));
    }

    public DataSystemColq(int device, Layout layout) {
        this.layout = layout;
        try (Arena a = Arena.ofConfined()) {
            MemorySegment out = a.allocate(ADDRESS);
            if ((int) colq_abi_version.invokeExact() != ABI_VERSION) throw new IllegalStateException("libcolq.so ABI version mismatch");
            int st = (int) colq_create.invokeExact(device, out);
            if (st != OK) throw new IllegalStateException("colq_create failed (" + st + "): no usable sm_100 GPU; libcolq has no CPU fallback");
            ctx = out.get(ADDRESS, 0);
        } catch (RuntimeException e) {
            throw e;
        } catch (Throwable t) {
            throw new IllegalStateException(t);
        }
    }

    /** DataSystemSerialIndices.register (:27-29): only records the reference; upload happens at the first execute. */
    public synchronized void register(String tableName, Table table) {
        tables.put(tableName, table);
    }

    @Override
    public synchronized QueryResult execute(Query query) {
        Objects.requireNonNull(query, "The 'query' argument must not be null");
        if (!tables.containsKey(query.tableName)) {
            return new QueryResult.Failure("The query targets the table '%s' but that table is not registered".formatted(query.tableName));
        }
        Table table = tables.get(query.tableName);
        try (Arena call = Arena.ofConfined()) {
            syncTables(call);
            MemorySegment qOut = call.allocate(ADDRESS);
            check((int) colq_query_create.invokeExact(ctx, call.allocateFrom(query.tableName), qOut));
            MemorySegment q = qOut.get(ADDRESS, 0);
            try {
                String opaque = translate(call, q, query, table);
                if (opaque != null) return new QueryResult.Failure(opaque);
                long words = (table.size() + 63L) / 64L;
                MemorySegment mask = call.allocate(JAVA_LONG, Math.max(words, 1));
                MemorySegment count = call.allocate(JAVA_LONG);
                int st = (int) colq_execute.invokeExact(ctx, q, mask, words, MemorySegment.NULL, 0L, count, MemorySegment.NULL);
                if (st == FAILURE) return new QueryResult.Failure(lastError());
                check(st);
                // BitSet.valueOf(LongBuffer): libcolq writes java.util.BitSet words directly
                BitSet matchingRows = BitSet.valueOf(mask.asSlice(0, words * 8).asByteBuffer().order(ByteOrder.LITTLE_ENDIAN).asLongBuffer());
                return new QueryResult.Success(table.subset(matchingRows));   // DataSystemSerialIndices.java:100
            } finally {
                int ignored = (int) colq_query_destroy.invokeExact(q);
            }
        } catch (RuntimeException e) {
            throw e;
        } catch (Throwable t) {
            throw new IllegalStateException(t);
        }
    }

    // ------------------------------------------------------------------------------------------ Query -> colq_query
    private String translate(Arena call, MemorySegment q, Query query, Table root) throws Throwable {
        record Pending(Query.Node node, int id, Table table) {}
        var stack = new ArrayDeque<Pending>();
        stack.push(new Pending(query.rootNode, 0, root));
        while (!stack.isEmpty()) {
            Pending p = stack.pop();
            for (Criteria c : p.node().getCriteria()) {
                switch (c) {
                    case Criteria.IntCriteria(int ordinal, var pred) -> {
                        int[] distinctInts = p.table() == null ? null : intDictionaryValues.getOrDefault(p.table(), Map.of()).get(ordinal);
                        if (!(pred instanceof Predicates.IntRange) && distinctInts != null) {
                            // an opaque IntPredicate (Runner.java:231): run it once per DISTINCT value, ship the accept set
                            MemorySegment accept = call.allocate(JAVA_LONG, distinctInts.length / 64 + 1);
                            for (int d = 0; d < distinctInts.length; d++)
                                if (pred.test(distinctInts[d])) accept.setAtIndex(JAVA_LONG, d >> 6, accept.getAtIndex(JAVA_LONG, d >> 6) | (1L << (d & 63)));
                            check((int) colq_query_criteria_i32_accept.invokeExact(q, p.id(), ordinal, accept, (long) distinctInts.length));
                            continue;
                        }
                        if (!(pred instanceof Predicates.IntRange(int lo, int hi)))
                            return "The criterion on ordinal %d is an opaque IntPredicate lambda; the GPU engine only runs structured predicates (dgroomes.data_system_b200.Predicates) and has no CPU fallback.".formatted(ordinal);
                        check((int) colq_query_criteria_i32_range.invokeExact(q, p.id(), ordinal, lo, hi));
                    }
                    case Criteria.StringCriteria(int ordinal, var pred) -> {
                        String[] distinct = p.table() == null ? null : dictionaryValues.getOrDefault(p.table(), Map.of()).get(ordinal);
                        if (!(pred instanceof Predicates.StringOp) && distinct != null) {
                            // an opaque Predicate<String>: run it once per DISTINCT value, ship the accept set
                            long words = distinct.length / 64 + 1;
                            MemorySegment accept = call.allocate(JAVA_LONG, words);
                            for (int d = 0; d < distinct.length; d++)
                                if (pred.test(distinct[d])) accept.setAtIndex(JAVA_LONG, d >> 6, accept.getAtIndex(JAVA_LONG, d >> 6) | (1L << (d & 63)));
                            check((int) colq_query_criteria_str_accept.invokeExact(q, p.id(), ordinal, accept, (long) distinct.length));
                            continue;
                        }
                        if (!(pred instanceof Predicates.StringOp op))
                            return "The criterion on ordinal %d is an opaque Predicate<String> lambda; the GPU engine only runs structured predicates (dgroomes.data_system_b200.Predicates) and has no CPU fallback.".formatted(ordinal);
                        byte[] needle = op.needle();
                        MemorySegment n = call.allocate(Math.max(needle.length, 1));
                        MemorySegment.copy(needle, 0, n, JAVA_BYTE, 0, needle.length);
                        check((int) colq_query_criteria_str.invokeExact(q, p.id(), ordinal, op.op().ordinal(), n, needle.length));
                    }
                }
            }
            for (Map.Entry<Integer, Query.Node> e : p.node().getChildrenByOrdinal().entrySet()) {
                MemorySegment out = call.allocate(JAVA_INT);
                check((int) colq_query_child.invokeExact(q, p.id(), (int) e.getKey(), out));
                Table child = null;
                if (p.table() != null && e.getKey() >= 0 && e.getKey() < p.table().columns().size()
                        && p.table().columns().get(e.getKey()) instanceof AssociationColumn ac) child = ac.associatedEntity();
                stack.push(new Pending(e.getValue(), out.get(JAVA_INT, 0), child));
            }
        }
        return null;
    }

    // ------------------------------------------------------------------------------------------ Table graph -> HBM
    private void syncTables(Arena call) throws Throwable {
        // discover every table reachable through association columns (identity-keyed, cycle-safe): the app registers
        // tables BEFORE associateTo appends their association columns (Runner.java:107 vs :138,165,195)
        var seen = new IdentityHashMap<Table, Boolean>();
        var todo = new ArrayDeque<>(tables.values());
        while (!todo.isEmpty()) {
            Table t = todo.pop();
            if (seen.put(t, true) != null) continue;
            for (Column c : t.columns()) if (c instanceof AssociationColumn ac) todo.push(ac.associatedEntity());
        }
        for (Table t : seen.keySet()) {
            if (handles.containsKey(t)) continue;
            MemorySegment out = call.allocate(JAVA_INT);
            check((int) colq_table_create.invokeExact(ctx, (long) t.size(), REPLICATED, 0L, out));
            handles.put(t, out.get(JAVA_INT, 0));
            uploadedColumns.put(t, 0);
        }
        for (Table t : seen.keySet()) {   // scalar columns
            int h = handles.get(t);
            List<? extends Column> cols = t.columns();
            for (int ordinal = uploadedColumns.get(t); ordinal < cols.size(); ordinal++) {
                switch (cols.get(ordinal)) {
                    case InMemoryColumn.IntegerColumn(int[] ints) when layout.dictionary() -> {
                        int[] distinct = java.util.Arrays.stream(ints).distinct().sorted().toArray();
                        int[] codes = new int[ints.length];
                        for (int i = 0; i < ints.length; i++) codes[i] = java.util.Arrays.binarySearch(distinct, ints[i]);
                        intDictionaryValues.computeIfAbsent(t, k -> new HashMap<>()).put(ordinal, distinct);
                        MemorySegment dict = call.allocate(JAVA_INT, Math.max(distinct.length, 1));
                        MemorySegment.copy(distinct, 0, dict, JAVA_INT, 0, distinct.length);
                        if (layout.hostResident() && codes.length > 0) {
                            MemorySegment seg = pinned(4L * codes.length);
                            MemorySegment.copy(codes, 0, seg, JAVA_INT, 0, codes.length);
                            check((int) colq_col_i32_dict_host.invokeExact(ctx, h, ordinal, seg, seg.byteSize(), (long) codes.length, dict, (long) distinct.length));
                        } else {
                            MemorySegment seg = call.allocate(JAVA_INT, Math.max(codes.length, 1));
                            MemorySegment.copy(codes, 0, seg, JAVA_INT, 0, codes.length);
                            check((int) colq_col_i32_dict.invokeExact(ctx, h, ordinal, seg, (long) codes.length, dict, (long) distinct.length));
                        }
                    }
                    case InMemoryColumn.IntegerColumn(int[] ints) -> {
                        if (layout.hostResident() && ints.length > 0) {
                            MemorySegment seg = pinned(4L * ints.length);
                            MemorySegment.copy(ints, 0, seg, JAVA_INT, 0, ints.length);
                            check((int) colq_col_i32_host.invokeExact(ctx, h, ordinal, seg, seg.byteSize(), (long) ints.length));
                        } else {
                            MemorySegment seg = call.allocate(JAVA_INT, Math.max(ints.length, 1));
                            MemorySegment.copy(ints, 0, seg, JAVA_INT, 0, ints.length);
                            check((int) colq_col_i32.invokeExact(ctx, h, ordinal, seg, (long) ints.length));
                        }
                    }
                    case InMemoryColumn.StringColumn(String[] strings) when layout.dictionary() && layout.deviceIngest() -> {
                        // plain offsets + UTF-8 bytes go up once; the GPU builds the dictionary (first-appearance order) and
                        // the distinct values come back for the opaque-lambda path
                        byte[][] enc = new byte[strings.length][];
                        long total = 0;
                        for (int i = 0; i < strings.length; i++) { enc[i] = strings[i].getBytes(StandardCharsets.UTF_8); total += enc[i].length; }
                        MemorySegment off = call.allocate(JAVA_INT, strings.length + 1L);
                        MemorySegment bytes = call.allocate(Math.max(total, 1));
                        long pos = 0;
                        for (int i = 0; i < strings.length; i++) {
                            off.setAtIndex(JAVA_INT, i, (int) pos);
                            MemorySegment.copy(enc[i], 0, bytes, JAVA_BYTE, pos, enc[i].length);
                            pos += enc[i].length;
                        }
                        off.setAtIndex(JAVA_INT, strings.length, (int) pos);
                        check((int) colq_col_str.invokeExact(ctx, h, ordinal, off, bytes, (long) strings.length, total));
                        MemorySegment nDict = call.allocate(JAVA_LONG), nBytes = call.allocate(JAVA_LONG);
                        check((int) colq_col_str_encode.invokeExact(ctx, h, ordinal, nDict));
                        int sizeQuery = (int) colq_col_dict_str.invokeExact(ctx, h, ordinal, MemorySegment.NULL, 0L, MemorySegment.NULL, 0L, nDict, nBytes);
                        if (sizeQuery != OK && sizeQuery != ERR_CAPACITY) check(sizeQuery);
                        long d = nDict.get(JAVA_LONG, 0), db = nBytes.get(JAVA_LONG, 0);
                        MemorySegment dOff = call.allocate(JAVA_INT, d + 1), dBytes = call.allocate(Math.max(db, 1));
                        check((int) colq_col_dict_str.invokeExact(ctx, h, ordinal, dOff, d + 1, dBytes, db, nDict, nBytes));
                        String[] distinct = new String[(int) d];
                        for (int i = 0; i < d; i++) {
                            int a = dOff.getAtIndex(JAVA_INT, i), b = dOff.getAtIndex(JAVA_INT, i + 1);
                            distinct[i] = new String(dBytes.asSlice(a, b - a).toArray(JAVA_BYTE), StandardCharsets.UTF_8);
                        }
                        dictionaryValues.computeIfAbsent(t, k -> new HashMap<>()).put(ordinal, distinct);
                    }
                    case InMemoryColumn.StringColumn(String[] strings) when layout.dictionary() -> {
                        // dictionary-encode while copying off-heap: distinct values in first-appearance order
                        var index = new HashMap<String, Integer>();
                        var distinct = new java.util.ArrayList<String>();
                        int[] codes = new int[strings.length];
                        for (int i = 0; i < strings.length; i++) {
                            Integer code = index.get(strings[i]);
                            if (code == null) { code = distinct.size(); index.put(strings[i], code); distinct.add(strings[i]); }
                            codes[i] = code;
                        }
                        dictionaryValues.computeIfAbsent(t, k -> new HashMap<>()).put(ordinal, distinct.toArray(String[]::new));
                        byte[][] enc = new byte[distinct.size()][];
                        long total = 0;
                        for (int i = 0; i < enc.length; i++) { enc[i] = distinct.get(i).getBytes(StandardCharsets.UTF_8); total += enc[i].length; }
                        MemorySegment off = call.allocate(JAVA_INT, enc.length + 1L);
                        MemorySegment bytes = call.allocate(Math.max(total, 1));
                        long pos = 0;
                        for (int i = 0; i < enc.length; i++) {
                            off.setAtIndex(JAVA_INT, i, (int) pos);
                            MemorySegment.copy(enc[i], 0, bytes, JAVA_BYTE, pos, enc[i].length);
                            pos += enc[i].length;
                        }
                        off.setAtIndex(JAVA_INT, enc.length, (int) pos);
                        if (layout.hostResident() && codes.length > 0) {
                            MemorySegment seg = pinned(4L * codes.length);
                            MemorySegment.copy(codes, 0, seg, JAVA_INT, 0, codes.length);
                            check((int) colq_col_str_dict_host.invokeExact(ctx, h, ordinal, seg, seg.byteSize(), (long) codes.length, off, bytes, (long) enc.length, total));
                        } else {
                            MemorySegment seg = call.allocate(JAVA_INT, Math.max(codes.length, 1));
                            MemorySegment.copy(codes, 0, seg, JAVA_INT, 0, codes.length);
                            check((int) colq_col_str_dict.invokeExact(ctx, h, ordinal, seg, (long) codes.length, off, bytes, (long) enc.length, total));
                        }
                    }
                    case InMemoryColumn.StringColumn(String[] strings) -> {
                        // offsets + UTF-8 bytes: byte equality == String.equals, byte substring == String.contains
                        byte[][] enc = new byte[strings.length][];
                        long total = 0;
                        for (int i = 0; i < strings.length; i++) { enc[i] = strings[i].getBytes(StandardCharsets.UTF_8); total += enc[i].length; }
                        boolean host = layout.hostResident() && strings.length > 0;
                        MemorySegment off = host ? pinned(4L * (strings.length + 1L)) : call.allocate(JAVA_INT, strings.length + 1L);
                        MemorySegment bytes = host ? pinned(total) : call.allocate(Math.max(total, 1));
                        long pos = 0;
                        for (int i = 0; i < strings.length; i++) {
                            off.setAtIndex(JAVA_INT, i, (int) pos);
                            MemorySegment.copy(enc[i], 0, bytes, JAVA_BYTE, pos, enc[i].length);
                            pos += enc[i].length;
                        }
                        off.setAtIndex(JAVA_INT, strings.length, (int) pos);
                        if (host) check((int) colq_col_str_host.invokeExact(ctx, h, ordinal, off, off.byteSize(), bytes, bytes.byteSize(), (long) strings.length, total));
                        else check((int) colq_col_str.invokeExact(ctx, h, ordinal, off, bytes, (long) strings.length, total));
                    }
                    case InMemoryColumn.BooleanColumn(boolean[] bools) -> {
                        MemorySegment seg = call.allocate(Math.max(bools.length, 1));
                        for (int i = 0; i < bools.length; i++) seg.set(JAVA_BYTE, i, (byte) (bools[i] ? 1 : 0));
                        check((int) colq_col_bool.invokeExact(ctx, h, ordinal, seg, (long) bools.length));
                    }
                    default -> { /* association columns below */ }
                }
            }
        }
        for (Table t : seen.keySet()) {   // association pairs, declared from their to-one (or smaller) side
            int h = handles.get(t);
            List<? extends Column> cols = t.columns();
            for (int ordinal = uploadedColumns.get(t); ordinal < cols.size(); ordinal++) {
                if (!(cols.get(ordinal) instanceof InMemoryColumn.AssociationColumn ac)) continue;
                var rev = ac.reverseAssociatedColumn();
                Table y = ac.associatedEntity();
                int yOrdinal = identityIndexOf(y.columns(), rev);
                // each pair is uploaded once: from the side that comes first in (table handle, ordinal) order
                int hy = handles.get(y);
                if (hy < h || (hy == h && yOrdinal < ordinal)) continue;
                uploadAssociation(call, h, ordinal, hy, yOrdinal, ac.associations);
            }
        }
        for (Table t : seen.keySet()) uploadedColumns.put(t, t.columns().size());
        for (var e : tables.entrySet()) {
            int h = handles.get(e.getValue());
            if (!Integer.valueOf(h).equals(registeredHandle.get(e.getKey()))) {
                check((int) colq_register.invokeExact(ctx, call.allocateFrom(e.getKey()), h));
                registeredHandle.put(e.getKey(), h);
            }
        }
    }

    /** Association[] (None | One | Many, Association.java:27-51) -> dense fk (-1 = None) or CSR. */
    private void uploadAssociation(Arena call, int x, int xOrdinal, int y, int yOrdinal, Association[] assoc) throws Throwable {
        boolean toOne = true;
        long nnz = 0;
        for (Association a : assoc) {
            if (a instanceof Association.Many(int[] idx)) { toOne = false; nnz += idx.length; }
            else if (a instanceof Association.One) nnz++;
        }
        if (layout.deviceIngest()) toOne = false;   // always ship the CSR: the GPU validates it and picks the representation
        if (toOne) {
            boolean host = layout.hostResident() && assoc.length > 0;
            MemorySegment fk = host ? pinned(4L * assoc.length) : call.allocate(JAVA_INT, Math.max(assoc.length, 1));
            for (int i = 0; i < assoc.length; i++) fk.setAtIndex(JAVA_INT, i, assoc[i] instanceof Association.One(int idx) ? idx : -1);
            if (host) check((int) colq_associate_fk_host.invokeExact(ctx, x, xOrdinal, y, yOrdinal, fk, fk.byteSize(), (long) assoc.length));
            else check((int) colq_associate_fk.invokeExact(ctx, x, xOrdinal, y, yOrdinal, fk, (long) assoc.length));
        } else {
            MemorySegment off = call.allocate(JAVA_LONG, assoc.length + 1L);
            MemorySegment tgt = call.allocate(JAVA_INT, Math.max(nnz, 1));
            long pos = 0;
            for (int i = 0; i < assoc.length; i++) {
                off.setAtIndex(JAVA_LONG, i, pos);
                switch (assoc[i]) {
                    case Association.Many(int[] idx) -> { for (int v : idx) tgt.setAtIndex(JAVA_INT, pos++, v); }
                    case Association.One(int idx) -> tgt.setAtIndex(JAVA_INT, pos++, idx);
                    case Association.None ignored -> { }
                }
            }
            off.setAtIndex(JAVA_LONG, assoc.length, pos);
            if (layout.deviceIngest()) check((int) colq_associate.invokeExact(ctx, x, xOrdinal, y, yOrdinal, off, tgt, (long) assoc.length, nnz, MemorySegment.NULL));
            else check((int) colq_associate_csr.invokeExact(ctx, x, xOrdinal, y, yOrdinal, off, tgt, (long) assoc.length, nnz));
        }
    }

    /** A pinned, device-mapped off-heap segment of at least {@code bytes} bytes, padded for whole-line reads. */
    private MemorySegment pinned(long bytes) throws Throwable {
        long padded = (bytes + 15) / 16 * 16 + 64;
        try (Arena a = Arena.ofConfined()) {
            MemorySegment out = a.allocate(ADDRESS);
            check((int) colq_host_alloc.invokeExact(ctx, padded, out));
            MemorySegment seg = out.get(ADDRESS, 0).reinterpret(padded);
            seg.fill((byte) 0);
            hostBuffers.add(seg);
            return seg;
        }
    }

    private static int identityIndexOf(List<? extends Column> cols, Object col) {
        for (int i = 0; i < cols.size(); i++) if (cols.get(i) == col) return i;
        throw new IllegalStateException("reverse association column is not a column of its table");
    }

    private String lastError() throws Throwable {
        MemorySegment p = (MemorySegment) colq_last_error.invokeExact(ctx);
        return p.reinterpret(1024).getString(0);
    }

    /** colq_status -> the exception class the reference would throw (include/colq.h). */
    private void check(int status) throws Throwable {
        switch (status) {
            case OK -> { }
            case THROW_INDEX_OOB -> throw new IndexOutOfBoundsException(lastError());
            case THROW_NULL -> throw new NullPointerException(lastError());
            case THROW_ILLEGAL_ARG -> throw new IllegalArgumentException(lastError());
            default -> throw new IllegalStateException("libcolq status " + status + ": " + lastError());
        }
    }

    /** Forget a registered table and release its device memory (the reference leaves this to the garbage collector). */
    public synchronized void unregister(String tableName) {
        Table t = tables.remove(tableName);
        registeredHandle.remove(tableName);
        if (t == null || tables.containsValue(t)) return;
        Integer h = handles.remove(t);
        uploadedColumns.remove(t);
        dictionaryValues.remove(t);
        intDictionaryValues.remove(t);
        try {
            if (h != null) check((int) colq_table_destroy.invokeExact(ctx, (int) h));
        } catch (RuntimeException e) {
            throw e;
        } catch (Throwable e) {
            throw new IllegalStateException(e);
        }
    }

    @Override
    public synchronized void close() {
        try {
            for (MemorySegment seg : hostBuffers) { int ignored = (int) colq_host_free.invokeExact(ctx, seg); }
            hostBuffers.clear();
            int ignored = (int) colq_destroy.invokeExact(ctx);
        } catch (Throwable t) {
            throw new IllegalStateException(t);
        }
        arena.close();
    }
}
