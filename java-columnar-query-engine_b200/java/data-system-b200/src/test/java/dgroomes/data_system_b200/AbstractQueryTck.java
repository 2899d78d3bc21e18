package dgroomes.data_system_b200;

import dgroomes.data_system.Association;
import dgroomes.data_system.Criteria;
import dgroomes.data_system.DataSystem;
import dgroomes.data_system.Query;
import dgroomes.data_system.QueryResult;
import dgroomes.data_system.Table;
import dgroomes.in_memory.InMemoryColumn;
import org.junit.jupiter.api.AfterEach;
import org.junit.jupiter.api.BeforeEach;
import org.junit.jupiter.api.Test;

import java.util.List;

import static dgroomes.in_memory.InMemoryColumn.ofInts;
import static dgroomes.in_memory.InMemoryColumn.ofStrings;
import static dgroomes.in_memory.InMemoryTable.ofColumns;
import static org.assertj.core.api.Assertions.assertThat;
import static org.assertj.core.api.Assertions.assertThatThrownBy;

/**
 * The Test Compatibility Kit the reference wishes for (README.md:149-153): "one functional test suite run against every
 * DataSystem implementation".  The five cases of the reference's own QueryTest
 * (data-system-serial-indices-arrays/src/test/java/dgroomes/queryengine/QueryTest.java:37,78,113,150,231) with the lambdas
 * replaced by structured predicates that implement the same functional interfaces (so the SAME test body also runs on the
 * reference's serial engine, see SerialIndicesTckTest), the five failure paths of Verifier.java /
 * DataSystemSerialIndices.java that the reference never asserts, and two association shapes its tests do not reach.
 * <p>
 * The identical cases run in this repository through the ctypes twin of the shim (tests/tck.py: 12 cases x 9 physical
 * layouts on one GPU, and every table sharded over 2 and 8 GPUs); this JUnit form is for a machine with JDK 22 and a B200.
 */
abstract class AbstractQueryTck {

    /** A DataSystem plus its (non-interface) register method, DataSystemSerialIndices.java:27. */
    interface Engine extends AutoCloseable {
        DataSystem system();
        void register(String name, Table table);
        @Override void close();
    }

    abstract Engine newEngine();

    Engine engine;

    @BeforeEach
    void setUp() { engine = newEngine(); }

    @AfterEach
    void tearDown() { engine.close(); }

    private List<? extends dgroomes.data_system.Column> success(Query query) {
        QueryResult result = engine.system().execute(query);
        assertThat(result).isInstanceOf(QueryResult.Success.class);
        return ((QueryResult.Success) result).resultSet().columns();
    }

    private String failure(Query query) {
        QueryResult result = engine.system().execute(query);
        assertThat(result).isInstanceOf(QueryResult.Failure.class);
        return ((QueryResult.Failure) result).message();
    }

    // ---------------------------------------------------------------------------- the reference's five QueryTest cases
    @Test
    void intQuery_oneColumnTable() {   // QueryTest.java:37-73
        engine.register("ints", ofColumns(ofInts(-1, 0, 1, 2, 3)));
        var query = new Query("ints");
        query.rootNode.addCriteria(new Criteria.IntCriteria(0, Predicates.intGreaterThan(0)));
        var columns = success(query);
        assertThat(columns).hasSize(1);
        assertThat(((InMemoryColumn.IntegerColumn) columns.get(0)).ints()).containsExactly(1, 2, 3);
    }

    @Test
    void intQuery_twoColumnTable() {   // QueryTest.java:78-108
        engine.register("cities", ofColumns(ofStrings("Minneapolis", "Rochester", "Duluth"), ofInts(425_336, 121_395, 86_697)));
        var query = new Query("cities");
        query.rootNode.addCriteria(new Criteria.IntCriteria(1, Predicates.intBetweenExclusive(100_000, 150_000)));
        var columns = success(query);
        assertThat(columns).hasSize(2);
        assertThat(((InMemoryColumn.StringColumn) columns.get(0)).strings()).containsExactly("Rochester");
    }

    @Test
    void multiCriteria_rootEntity() {   // QueryTest.java:113-144
        engine.register("strings", ofColumns(ofStrings("a", "a", "b", "c", "c", "d")));
        var query = new Query("strings");
        query.rootNode.addCriteria(new Criteria.StringCriteria(0, Predicates.strCompareGt("a")))
                .addCriteria(new Criteria.StringCriteria(0, Predicates.strCompareLt("d")));
        assertThat(((InMemoryColumn.StringColumn) success(query).get(0)).strings()).containsExactly("b", "c", "c");
    }

    @Test
    void queryOnAssociationProperty() {   // QueryTest.java:150-229
        var cities = ofColumns(ofStrings("Minneapolis", "Pierre", "Duluth"));
        engine.register("cities", cities);
        var states = ofColumns(ofStrings("Minnesota", "South Dakota"));
        engine.register("states", states);
        cities.associateTo(states, Association.toOne(0), Association.toOne(1), Association.toOne(0));
        for (var c : List.of(List.of("South Dakota", List.of("Pierre")), List.of("Minnesota", List.of("Minneapolis", "Duluth")))) {
            var query = new Query("cities");
            query.rootNode.createChild(1).addCriteria(new Criteria.StringCriteria(0, Predicates.strEquals((String) c.get(0))));
            var columns = success(query);
            assertThat(columns).hasSize(2);
            assertThat(List.of(((InMemoryColumn.StringColumn) columns.get(0)).strings())).isEqualTo(c.get(1));
        }
    }

    @Test
    void multiCriteria_includingIntermediateEntity() {   // QueryTest.java:231-343
        var sections = ofColumns(
                ofStrings("maple trees", "lilacs", "", "", "", "", "Boston ferns", "rose bush", "cedar trees"),
                ofStrings("trees", "shrubs", "", "", "", "", "ferns", "shrubs", "trees"));
        engine.register("sections", sections);
        sections.associateTo(sections,
                Association.toMany(1, 3), Association.toMany(0, 2, 4), Association.toMany(1, 5),
                Association.toMany(0, 4, 6), Association.toMany(1, 3, 5, 7), Association.toMany(2, 4, 8),
                Association.toMany(3, 7), Association.toMany(4, 6, 8), Association.toMany(5, 7));
        var query = new Query("sections");
        query.rootNode.addCriteria(new Criteria.StringCriteria(1, Predicates.strEquals("trees")))
                .createChild(2).addCriteria(new Criteria.StringCriteria(1, Predicates.strEquals("shrubs")))
                .createChild(2).addCriteria(new Criteria.StringCriteria(1, Predicates.strEquals("ferns")));
        var columns = success(query);
        assertThat(columns).hasSize(4);
        assertThat(((InMemoryColumn.StringColumn) columns.get(0)).strings()).containsExactly("cedar trees");
    }

    // ---------------------------------------------------------------------------- failure paths (never asserted by the reference)
    @Test
    void failure_unregisteredTable() {   // DataSystemSerialIndices.java:54-57
        assertThat(failure(new Query("nope"))).isEqualTo("The query targets the table 'nope' but that table is not registered");
    }

    @Test
    void failure_typeMismatch() {   // Verifier.java:73-74, 78-79
        engine.register("t", ofColumns(ofStrings("a"), ofInts(1)));
        var q = new Query("t");
        q.rootNode.addCriteria(new Criteria.IntCriteria(0, Predicates.intRange(0, 1)));
        assertThat(failure(q)).isEqualTo("The column is a string column but the criterion is not a string predicate.");
        q = new Query("t");
        q.rootNode.addCriteria(new Criteria.StringCriteria(1, Predicates.strEquals("a")));
        assertThat(failure(q)).isEqualTo("The column is an integer column but the criterion is not an integer predicate.");
    }

    @Test
    void failure_ordinalOutOfBounds() {   // Verifier.java:62-67: `size() < ordinal`, so ordinal == width reaches columns().get()
        engine.register("t", ofColumns(ofInts(1, 2)));
        var q = new Query("t");
        q.rootNode.addCriteria(new Criteria.IntCriteria(5, Predicates.intRange(0, 1)));
        assertThat(failure(q)).isEqualTo("The query ordinal '5' is out of bounds for the table with 1 columns");
        var q2 = new Query("t");
        q2.rootNode.addCriteria(new Criteria.IntCriteria(1, Predicates.intRange(0, 1)));
        assertThatThrownBy(() -> engine.system().execute(q2)).isInstanceOf(IndexOutOfBoundsException.class);
    }

    @Test
    void failure_booleanAndAssociationCriteria() {   // Verifier.java:82-87
        var t = ofColumns(ofInts(1, 2), new InMemoryColumn.BooleanColumn(new boolean[]{true, false}));
        var u = ofColumns(ofInts(7));
        t.associateTo(u, Association.toOne(0), Association.toNone());
        engine.register("t", t);
        var q = new Query("t");
        q.rootNode.addCriteria(new Criteria.IntCriteria(1, Predicates.intRange(0, 1)));
        assertThat(failure(q)).isEqualTo("Boolean columns are not supported yet.");
        q = new Query("t");
        q.rootNode.addCriteria(new Criteria.IntCriteria(2, Predicates.intRange(0, 1)));
        assertThat(failure(q)).isEqualTo("Association columns can't be matched on with a scalar criteria.");
    }

    @Test
    void failure_childNotAssociation() {   // Verifier.java:100-104
        engine.register("t", ofColumns(ofInts(1, 2)));
        var q = new Query("t");
        q.rootNode.createChild(0);
        assertThat(failure(q)).startsWith("The column at ordinal 0 is not an association column.");
        var q2 = new Query("t");
        q2.rootNode.createChild(3);
        assertThatThrownBy(() -> engine.system().execute(q2)).isInstanceOf(IndexOutOfBoundsException.class);
    }

    // ---------------------------------------------------------------------------- shapes the reference's tests do not reach
    @Test
    void noneAndMixedReverse() {   // None rows, and a reverse column that mixes None / One / Many (InMemoryTable.java:55-82)
        var owners = ofColumns(ofStrings("ann", "bob", "cy", "dee"));
        var pets = ofColumns(ofStrings("rex", "tom", "kit", "jay", "moe"));
        engine.register("owners", owners);
        engine.register("pets", pets);
        pets.associateTo(owners, Association.toOne(0), Association.toOne(2), Association.toNone(), Association.toOne(0), Association.toOne(2));
        var q = new Query("owners");   // owners that have a pet named tom or moe, through the reverse column
        q.rootNode.createChild(1).addCriteria(new Criteria.StringCriteria(0, Predicates.strContains("o")));
        assertThat(((InMemoryColumn.StringColumn) success(q).get(0)).strings()).containsExactly("cy");
        var q2 = new Query("pets");    // pets whose owner's name contains "n"
        q2.rootNode.createChild(1).addCriteria(new Criteria.StringCriteria(0, Predicates.strContains("n")));
        assertThat(((InMemoryColumn.StringColumn) success(q2).get(0)).strings()).containsExactly("rex", "jay");
    }

    @Test
    void twoChildrenOnOneNode() {
        var people = ofColumns(ofStrings("p0", "p1", "p2", "p3"));
        var towns = ofColumns(ofStrings("north", "south"));
        var jobs = ofColumns(ofStrings("baker", "smith", "clerk"));
        people.associateTo(towns, Association.toOne(0), Association.toOne(1), Association.toOne(0), Association.toOne(1));
        people.associateTo(jobs, Association.toOne(0), Association.toOne(0), Association.toOne(1), Association.toMany(1, 2));
        engine.register("people", people);
        var q = new Query("people");
        q.rootNode.createChild(1).addCriteria(new Criteria.StringCriteria(0, Predicates.strEquals("south")));
        q.rootNode.createChild(2).addCriteria(new Criteria.StringCriteria(0, Predicates.strEquals("smith")));
        assertThat(((InMemoryColumn.StringColumn) success(q).get(0)).strings()).containsExactly("p3");
    }
}
