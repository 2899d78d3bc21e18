// tck_main.cpp -- the reference's QueryTest (data-system-serial-indices-arrays/src/test/java/dgroomes/queryengine/
// QueryTest.java) restated against the C++ host mirror; links libcolq.so only through its C ABI.  Needs a B200.
// Run by tests/test_gpu_cpp_tck.py.  Exit code 0 = every case passed.
#include <cstdio>
#include <functional>

#include "colq.hpp"

using namespace colq;

static int failures = 0;
static Options g_options;  // every case runs once per physical layout (device / host-resident x plain / dictionary)
#define EXPECT(cond)                                                          \
    do {                                                                      \
        if (!(cond)) { std::printf("  FAILED %s:%d: %s\n", __FILE__, __LINE__, #cond); ++failures; } \
    } while (0)

static std::shared_ptr<InMemoryTable> success(const QueryResult& r) {
    if (auto* f = std::get_if<Failure>(&r)) throw std::runtime_error("query failed: " + f->message);
    return std::get<Success>(r).resultSet;
}

static void intQuery_oneColumnTable() {  // QueryTest.java:37-73
    DataSystemColq ds(0, g_options);
    ds.registerTable("ints", InMemoryTable::ofColumns({ofInts({-1, 0, 1, 2, 3})}));
    Query query("ints");
    query.rootNode.addCriteria(IntCriteria{0, intGreaterThan(0)});
    const auto tp = success(ds.execute(query));
    const InMemoryTable& t = *tp;
    EXPECT(t.width() == 1);
    EXPECT((std::get<IntegerColumn>(t.columns[0]).ints == std::vector<int32_t>{1, 2, 3}));
}

static void intQuery_twoColumnTable() {  // QueryTest.java:78-108
    DataSystemColq ds(0, g_options);
    ds.registerTable("cities", InMemoryTable::ofColumns({ofStrings({"Minneapolis", "Rochester", "Duluth"}), ofInts({425336, 121395, 86697})}));
    Query query("cities");
    query.rootNode.addCriteria(IntCriteria{1, intBetweenExclusive(100000, 150000)});
    const auto tp = success(ds.execute(query));
    const InMemoryTable& t = *tp;
    EXPECT(t.width() == 2);
    EXPECT((std::get<StringColumn>(t.columns[0]).strings == std::vector<std::string>{"Rochester"}));
}

static void multiCriteria_rootEntity() {  // QueryTest.java:113-144
    DataSystemColq ds(0, g_options);
    ds.registerTable("strings", InMemoryTable::ofColumns({ofStrings({"a", "a", "b", "c", "c", "d"})}));
    Query query("strings");
    query.rootNode.addCriteria(StringCriteria{0, strCompareGt("a")}).addCriteria(StringCriteria{0, strCompareLt("d")});
    const auto tp = success(ds.execute(query));
    const InMemoryTable& t = *tp;
    EXPECT((std::get<StringColumn>(t.columns[0]).strings == std::vector<std::string>{"b", "c", "c"}));
}

static void queryOnAssociationProperty() {  // QueryTest.java:150-229
    DataSystemColq ds(0, g_options);
    auto cities = InMemoryTable::ofColumns({ofStrings({"Minneapolis", "Pierre", "Duluth"})});
    ds.registerTable("cities", cities);
    auto states = InMemoryTable::ofColumns({ofStrings({"Minnesota", "South Dakota"})});
    ds.registerTable("states", states);
    cities->associateTo(*states, {toOne(0), toOne(1), toOne(0)});
    {
        Query query("cities");
        query.rootNode.createChild(1).addCriteria(StringCriteria{0, strEquals("South Dakota")});
        const auto tp = success(ds.execute(query));
        const InMemoryTable& t = *tp;
        EXPECT(t.width() == 2);
        EXPECT((std::get<StringColumn>(t.columns[0]).strings == std::vector<std::string>{"Pierre"}));
    }
    {
        Query query("cities");
        query.rootNode.createChild(1).addCriteria(StringCriteria{0, strEquals("Minnesota")});
        const auto tp = success(ds.execute(query));
        const InMemoryTable& t = *tp;
        EXPECT((std::get<StringColumn>(t.columns[0]).strings == std::vector<std::string>{"Minneapolis", "Duluth"}));
    }
}

static void multiCriteria_includingIntermediateEntity() {  // QueryTest.java:231-343
    DataSystemColq ds(0, g_options);
    auto sections = InMemoryTable::ofColumns({
        ofStrings({"maple trees", "lilacs", "", "", "", "", "Boston ferns", "rose bush", "cedar trees"}),
        ofStrings({"trees", "shrubs", "", "", "", "", "ferns", "shrubs", "trees"})});
    ds.registerTable("sections", sections);
    sections->associateTo(*sections, {toMany({1, 3}), toMany({0, 2, 4}), toMany({1, 5}), toMany({0, 4, 6}), toMany({1, 3, 5, 7}),
                                      toMany({2, 4, 8}), toMany({3, 7}), toMany({4, 6, 8}), toMany({5, 7})});
    Query query("sections");
    query.rootNode.addCriteria(StringCriteria{1, strEquals("trees")})
        .createChild(2).addCriteria(StringCriteria{1, strEquals("shrubs")})
        .createChild(2).addCriteria(StringCriteria{1, strEquals("ferns")});
    const auto tp = success(ds.execute(query));
    const InMemoryTable& t = *tp;
    EXPECT(t.width() == 4);
    EXPECT((std::get<StringColumn>(t.columns[0]).strings == std::vector<std::string>{"cedar trees"}));
}

static void failures_followTheVerifier() {  // E/Verifier.java:62-104, E/DataSystemSerialIndices.java:54-57
    DataSystemColq ds(0, g_options);
    ds.registerTable("t", InMemoryTable::ofColumns({ofStrings({"a"}), ofInts({1})}));
    auto r = ds.execute(Query("nope"));
    EXPECT(std::get<Failure>(r).message == "The query targets the table 'nope' but that table is not registered");
    Query q("t");
    q.rootNode.addCriteria(IntCriteria{0, IntRange{0, 1}});
    EXPECT(std::get<Failure>(ds.execute(q)).message == "The column is a string column but the criterion is not a string predicate.");
    Query q2("t");
    q2.rootNode.createChild(1);
    EXPECT(std::get<Failure>(ds.execute(q2)).message ==
           "The column at ordinal 1 is not an association column. It is a dgroomes.in_memory.InMemoryColumn$IntegerColumn");
    Query q3("t");
    q3.rootNode.addCriteria(IntCriteria{2, IntRange{0, 1}});  // ordinal == width: the reference's off-by-one -> IndexOutOfBounds
    bool threw = false;
    try { ds.execute(q3); } catch (const std::out_of_range&) { threw = true; }
    EXPECT(threw);
}

// The reference's criteria are opaque lambdas (QueryTest.java:169,194; Runner.java:236,255-259).  Over dictionary-encoded
// columns they run unchanged (evaluated per distinct value on the host); over plain columns they are a Failure.
static void opaqueLambdas_onlyOverDictionaryColumns() {
    DataSystemColq ds(0, g_options);
    auto cities = InMemoryTable::ofColumns({ofStrings({"Minneapolis", "Pierre", "Duluth", "Pierre"})});
    ds.registerTable("cities", cities);
    auto states = InMemoryTable::ofColumns({ofStrings({"Minnesota", "South Dakota"})});
    ds.registerTable("states", states);
    cities->associateTo(*states, {toOne(0), toOne(1), toOne(0), toOne(1)});
    Query query("cities");
    query.rootNode.addCriteria(StringCriteria{0, StringLambda([](const std::string& s) { return s.size() == 6; })})
        .createChild(1).addCriteria(StringCriteria{0, StringLambda([](const std::string& s) { return s.find("South") != std::string::npos; })});
    const QueryResult r = ds.execute(query);
    if (g_options.dictionary) {
        const auto tp = success(r);
        EXPECT((std::get<StringColumn>(tp->columns[0]).strings == std::vector<std::string>{"Pierre", "Pierre"}));
    } else {
        EXPECT(std::holds_alternative<Failure>(r) && std::get<Failure>(r).message.find("no CPU fallback") != std::string::npos);
    }
}

// criteria over a BooleanColumn (declared, never reached by the reference: E/Verifier.java:82-84)
static void booleanCriteria() {
    DataSystemColq ds(0, g_options);
    auto t = InMemoryTable::ofColumns({ofInts({0, 1, 2, 3, 4}), BooleanColumn{{1, 0, 0, 1, 1}}});
    ds.registerTable("t", t);
    Query q("t");
    q.rootNode.addCriteria(BooleanCriteria{1, [](bool b) { return !b; }}).addCriteria(IntCriteria{0, IntRange{0, 1}});
    EXPECT((std::get<IntegerColumn>(success(ds.execute(q))->columns[0]).ints == std::vector<int32_t>{1}));
    Query all("t");
    all.rootNode.addCriteria(BooleanCriteria{1, [](bool) { return true; }});
    EXPECT(std::get<IntegerColumn>(success(ds.execute(all))->columns[0]).ints.size() == 5);
    Query bad("t");   // every criterion the reference can express on a boolean column keeps its Failure
    bad.rootNode.addCriteria(IntCriteria{1, IntRange{0, 1}});
    EXPECT(std::get<Failure>(ds.execute(bad)).message == "Boolean columns are not supported yet.");
}

int main() {
    const std::pair<const char*, std::function<void()>> cases[] = {
        {"intQuery_oneColumnTable", intQuery_oneColumnTable},
        {"intQuery_twoColumnTable", intQuery_twoColumnTable},
        {"multiCriteria_rootEntity", multiCriteria_rootEntity},
        {"queryOnAssociationProperty", queryOnAssociationProperty},
        {"multiCriteria_includingIntermediateEntity", multiCriteria_includingIntermediateEntity},
        {"failures_followTheVerifier", failures_followTheVerifier},
        {"opaqueLambdas_onlyOverDictionaryColumns", opaqueLambdas_onlyOverDictionaryColumns},
        {"booleanCriteria", booleanCriteria},
    };
    const std::pair<const char*, Options> layouts[] = {
        {"device", Options{Residency::Device, false}}, {"host-resident", Options{Residency::Host, false}},
        {"device+dictionary", Options{Residency::Device, true}}, {"host-resident+dictionary", Options{Residency::Host, true}},
    };
    for (auto& l : layouts) {
        g_options = l.second;
        for (auto& c : cases) {
            std::printf("[ RUN ] %s (%s)\n", c.first, l.first);
            try { c.second(); } catch (const std::exception& e) { std::printf("  EXCEPTION %s\n", e.what()); ++failures; }
        }
    }
    std::printf("%s (%d failure(s))\n", failures ? "FAILED" : "PASSED", failures);
    return failures ? 1 : 0;
}
