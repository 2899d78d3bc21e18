package dgroomes.data_system_b200;

import dgroomes.data_system.Criteria;
import dgroomes.data_system.DataSystem;
import dgroomes.data_system.Query;
import dgroomes.data_system.QueryResult;
import dgroomes.data_system.Table;
import dgroomes.data_system_serial_indices_arrays.DataSystemSerialIndices;
import org.junit.jupiter.api.Test;

import static dgroomes.in_memory.InMemoryColumn.ofInts;
import static dgroomes.in_memory.InMemoryTable.ofColumns;
import static org.assertj.core.api.Assertions.assertThat;

/**
 * The TCK (AbstractQueryTck) against every engine and physical layout: the B200 engine with columns copied to HBM, with
 * pinned host-resident + dictionary-encoded columns ({@code Layout.forUnchangedLambdas()}), with the load-time work done
 * by the GPU ({@code Layout.deviceIngest()}), sharded over every visible GPU from this one JVM (DataSystemColqGroup) --
 * and against the reference's own serial engine, which must pass the very same cases.
 */
final class QueryTckTest {

    private static AbstractQueryTck.Engine colq(DataSystemColq.Layout layout) {
        var ds = new DataSystemColq(0, layout);
        return new AbstractQueryTck.Engine() {
            public DataSystem system() { return ds; }
            public void register(String name, Table table) { ds.register(name, table); }
            public void close() { ds.close(); }
        };
    }

    static final class DefaultLayout extends AbstractQueryTck {
        @Override Engine newEngine() { return colq(new DataSystemColq.Layout(false, false)); }

        @Test
        void opaqueLambdaIsAFailureNotAFallback() {
            engine.register("ints", ofColumns(ofInts(1, 2, 3)));
            var query = new Query("ints");
            query.rootNode.addCriteria(new Criteria.IntCriteria(0, i -> i > 1));
            assertThat(engine.system().execute(query)).isInstanceOf(QueryResult.Failure.class);
        }
    }

    static final class HostResidentDictionary extends AbstractQueryTck {
        @Override Engine newEngine() { return colq(DataSystemColq.Layout.forUnchangedLambdas()); }

        @Test
        void theReferencesOwnLambdasRunPerDistinctValue() {   // Runner.java:231,236: opaque lambdas over dictionary-encoded columns
            engine.register("ints", ofColumns(ofInts(5, 10_050, 7, 10_050, 99_999)));
            var query = new Query("ints");
            query.rootNode.addCriteria(new Criteria.IntCriteria(0, i -> i >= 10_000 && i < 10_100));
            var result = (QueryResult.Success) engine.system().execute(query);
            assertThat(result.resultSet().size()).isEqualTo(2);
        }
    }

    static final class DeviceIngest extends AbstractQueryTck {
        @Override Engine newEngine() { return colq(DataSystemColq.Layout.deviceIngest()); }
    }

    /** Every table split over all visible GPUs, association keys global, driven from this one JVM. */
    static final class AllGpusOneJvm extends AbstractQueryTck {
        @Override Engine newEngine() {
            int n = Integer.getInteger("colq.gpus", 2);
            int[] devices = new int[n];
            for (int i = 0; i < n; i++) devices[i] = i;
            var ds = new DataSystemColqGroup(devices);
            return new Engine() {
                public DataSystem system() { return ds; }
                public void register(String name, Table table) { ds.register(name, table); }
                public void close() { ds.close(); }
            };
        }
    }

    /** The reference's own engine passes the same suite (structured predicates are ordinary IntPredicate / Predicate). */
    static final class ReferenceSerialIndices extends AbstractQueryTck {
        @Override Engine newEngine() {
            var ds = new DataSystemSerialIndices();
            return new Engine() {
                public DataSystem system() { return ds; }
                public void register(String name, Table table) { ds.register(name, table); }
                public void close() { }
            };
        }
    }
}
