package dgroomes.data_system_b200;

import dgroomes.data_system.Association;
import dgroomes.data_system.Criteria;
import dgroomes.data_system.Query;
import dgroomes.data_system.QueryResult;
import dgroomes.in_memory.InMemoryColumn;
import org.junit.jupiter.api.AfterEach;
import org.junit.jupiter.api.BeforeEach;
import org.junit.jupiter.api.Test;

import static dgroomes.in_memory.InMemoryColumn.ofInts;
import static dgroomes.in_memory.InMemoryColumn.ofStrings;
import static dgroomes.in_memory.InMemoryTable.ofColumns;
import static org.assertj.core.api.Assertions.assertThat;

/**
 * The reference's QueryTest (data-system-serial-indices-arrays/src/test/java/dgroomes/queryengine/QueryTest.java)
 * against DataSystemColq, lambdas replaced by structured predicates. The same cases run in this repository through
 * the ctypes twin (tests/tck.py); this JUnit class is for a machine that has JDK 22 and a B200.
 */
class QueryTckTest {

    DataSystemColq dataSystem;

    @BeforeEach
    void setUp() { dataSystem = new DataSystemColq(); }

    @AfterEach
    void tearDown() { dataSystem.close(); }

    @Test
    void intQuery_oneColumnTable() {   // QueryTest.java:37-73
        dataSystem.register("ints", ofColumns(ofInts(-1, 0, 1, 2, 3)));
        var query = new Query("ints");
        query.rootNode.addCriteria(new Criteria.IntCriteria(0, Predicates.intGreaterThan(0)));
        var result = (QueryResult.Success) dataSystem.execute(query);
        assertThat(((InMemoryColumn.IntegerColumn) result.resultSet().columns().get(0)).ints()).containsExactly(1, 2, 3);
    }

    @Test
    void queryOnAssociationProperty() {   // QueryTest.java:150-229
        var cities = ofColumns(ofStrings("Minneapolis", "Pierre", "Duluth"));
        dataSystem.register("cities", cities);
        var states = ofColumns(ofStrings("Minnesota", "South Dakota"));
        dataSystem.register("states", states);
        cities.associateTo(states, Association.toOne(0), Association.toOne(1), Association.toOne(0));
        var query = new Query("cities");
        query.rootNode.createChild(1).addCriteria(new Criteria.StringCriteria(0, Predicates.strEquals("Minnesota")));
        var result = (QueryResult.Success) dataSystem.execute(query);
        assertThat(((InMemoryColumn.StringColumn) result.resultSet().columns().getFirst()).strings()).containsExactly("Minneapolis", "Duluth");
    }

    @Test
    void opaqueLambdaIsAFailure() {
        dataSystem.register("ints", ofColumns(ofInts(1, 2, 3)));
        var query = new Query("ints");
        query.rootNode.addCriteria(new Criteria.IntCriteria(0, i -> i > 1));
        assertThat(dataSystem.execute(query)).isInstanceOf(QueryResult.Failure.class);
    }
}
