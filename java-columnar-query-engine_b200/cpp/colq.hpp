// colq.hpp -- header-only C++17 mirror of the reference's data-system API over the C ABI of libcolq.so.
//
// The reference host language is Java (no JVM in this image); per the build rules the host side above the C ABI is
// written in C++ for compiled references.  Same names, argument meaning and error behaviour as
//   DS = data-system/src/main/java/dgroomes/data_system, M = data-model-in-memory/src/main/java/dgroomes/in_memory,
//   E  = data-system-serial-indices-arrays/src/main/java/dgroomes/data_system_serial_indices_arrays
// so that cpp/tck_main.cpp reads like the reference's QueryTest.  No CPU fallback: every query runs in libcolq.so.
#pragma once

#include <cstdint>
#include <cstring>
#include <functional>
#include <map>
#include <memory>
#include <stdexcept>
#include <string>
#include <utility>
#include <variant>
#include <vector>

#include "../../include/colq.h"

namespace colq {

// ---- DS/Association.java:6-52
struct None {};
struct One { int idx; };
struct Many { std::vector<int> indices; };
using Association = std::variant<None, One, Many>;
inline Association toNone() { return None{}; }
inline Association toOne(int i) { return One{i}; }
inline Association toMany(std::initializer_list<int> l) { return Many{std::vector<int>(l)}; }
inline Association add(const Association& a, int idx) {  // Association.add (:32-50)
    if (std::holds_alternative<None>(a)) return One{idx};
    if (auto* o = std::get_if<One>(&a)) return Many{{o->idx, idx}};
    Many m = std::get<Many>(a);
    m.indices.push_back(idx);
    return m;
}

// ---- java.util.BitSet word layout
struct BitSet {
    std::vector<uint64_t> words;
    bool get(int64_t i) const { return (words[i >> 6] >> (i & 63)) & 1u; }
    int64_t cardinality() const {
        int64_t c = 0;
        for (uint64_t w : words) c += __builtin_popcountll(w);
        return c;
    }
};

class InMemoryTable;

// ---- M/InMemoryColumn.java:19-138
struct IntegerColumn { std::vector<int32_t> ints; };
struct StringColumn { std::vector<std::string> strings; };
struct BooleanColumn { std::vector<uint8_t> bools; };
struct AssociationColumn {
    InMemoryTable* associatedEntity = nullptr;
    std::vector<Association> associations;
    InMemoryTable* owner = nullptr;
    int reverseOrdinal = -1;  // ordinal of reverseAssociatedColumn() in associatedEntity
    bool forward = true;
};
using InMemoryColumn = std::variant<IntegerColumn, StringColumn, BooleanColumn, AssociationColumn>;

inline InMemoryColumn ofInts(std::initializer_list<int32_t> v) { return IntegerColumn{std::vector<int32_t>(v)}; }
inline InMemoryColumn ofStrings(std::initializer_list<const char*> v) {
    StringColumn c;
    for (const char* s : v) c.strings.emplace_back(s);
    return c;
}

// ---- M/InMemoryTable.java:16-160
class InMemoryTable {
public:
    std::vector<InMemoryColumn> columns;

    static std::shared_ptr<InMemoryTable> ofColumns(std::initializer_list<InMemoryColumn> cols) {
        auto t = std::make_shared<InMemoryTable>();
        t->columns.assign(cols.begin(), cols.end());
        return t;
    }
    int width() const { return (int)columns.size(); }
    int64_t size() const {  // length of column 0 (:92-101)
        return std::visit([](auto&& c) -> int64_t {
            using T = std::decay_t<decltype(c)>;
            if constexpr (std::is_same_v<T, IntegerColumn>) return (int64_t)c.ints.size();
            else if constexpr (std::is_same_v<T, StringColumn>) return (int64_t)c.strings.size();
            else if constexpr (std::is_same_v<T, BooleanColumn>) return (int64_t)c.bools.size();
            else return (int64_t)c.associations.size();
        }, columns.at(0));
    }
    // x.associateTo(y, associations) (:44-90): appends the forward column here and the transposed column to y
    int associateTo(InMemoryTable& y, std::vector<Association> associations) {
        const int64_t ysize = y.size();
        std::vector<std::vector<int>> yToX((size_t)ysize);
        for (int x = 0; x < (int)associations.size(); ++x) {
            auto visit = [&](int t) {
                if (t < 0 || t >= ysize) throw std::out_of_range("NullPointerException: association target outside the associated table");
                yToX[(size_t)t].push_back(x);  // x ascending (:61)
            };
            if (auto* o = std::get_if<One>(&associations[x])) visit(o->idx);
            else if (auto* m = std::get_if<Many>(&associations[x])) for (int t : m->indices) visit(t);
        }
        AssociationColumn f;
        f.associatedEntity = &y; f.associations = std::move(associations); f.owner = this; f.forward = true;
        columns.emplace_back(std::move(f));
        const int xo = width() - 1;
        AssociationColumn r;
        r.associatedEntity = this; r.owner = &y; r.forward = false; r.reverseOrdinal = xo;
        for (auto& xs : yToX) {  // None / One / Many by list length (:75-82)
            if (xs.empty()) r.associations.emplace_back(None{});
            else if (xs.size() == 1) r.associations.emplace_back(One{xs[0]});
            else r.associations.emplace_back(Many{xs});
        }
        y.columns.emplace_back(std::move(r));
        std::get<AssociationColumn>(columns[(size_t)xo]).reverseOrdinal = y.width() - 1;
        return xo;
    }
    // M/InMemoryTable.java:106-159: every column pruned to the set bits, ascending, association indices un-remapped
    std::shared_ptr<InMemoryTable> subset(const BitSet& rows) const {
        auto out = std::make_shared<InMemoryTable>();
        const int64_t n = size();
        for (const auto& col : columns) {
            out->columns.push_back(std::visit([&](auto&& c) -> InMemoryColumn {
                using T = std::decay_t<decltype(c)>;
                T p = c;
                auto prune = [&](auto& vec) {
                    std::decay_t<decltype(vec)> keep;
                    for (int64_t i = 0; i < n; ++i) if (rows.get(i)) keep.push_back(vec[(size_t)i]);
                    vec = std::move(keep);
                };
                if constexpr (std::is_same_v<T, IntegerColumn>) prune(p.ints);
                else if constexpr (std::is_same_v<T, StringColumn>) prune(p.strings);
                else if constexpr (std::is_same_v<T, BooleanColumn>) prune(p.bools);
                else { prune(p.associations); p.reverseOrdinal = -1; }
                return p;
            }, col));
        }
        return out;
    }
};

// ---- structured predicates (stand-ins for the reference's lambdas, DS/Criteria.java:17-19)
struct IntRange { int32_t lo, hi; };
struct StringOp { colq_str_op op; std::string value; };
inline IntRange intGreaterThan(int32_t x) { return {x + 1, INT32_MAX}; }
inline IntRange intBetweenExclusive(int32_t lo, int32_t hi) { return {lo + 1, hi - 1}; }
inline IntRange intHalfOpen(int32_t lo, int32_t hi) { return {lo, hi - 1}; }
inline StringOp strEquals(std::string v) { return {COLQ_STR_EQ, std::move(v)}; }
inline StringOp strContains(std::string v) { return {COLQ_STR_CONTAINS, std::move(v)}; }
inline StringOp strCompareGt(std::string v) { return {COLQ_STR_CMP_GT, std::move(v)}; }
inline StringOp strCompareLt(std::string v) { return {COLQ_STR_CMP_LT, std::move(v)}; }

// ---- DS/Criteria.java:10-20, DS/Query.java:17-54, DS/QueryResult.java:3-9
// an opaque Predicate<String> like the reference's own lambdas (app/.../Runner.java:236,255-259): runs only over
// dictionary-encoded columns (Options::dictionary), evaluated once per DISTINCT value on the host
using StringLambda = std::function<bool(const std::string&)>;
struct IntCriteria { int ordinal; IntRange integerPredicate; };
struct StringCriteria { int ordinal; std::variant<StringOp, StringLambda> stringPredicate; };
// SURVEY.md 8(f4) extension: BooleanColumnFilterable.where(Predicate<Boolean>) (DS/ColumnFilterable.java:20-22), which the
// reference declares and its Verifier refuses (E/Verifier.java:82-84); evaluated on false and true by the host
struct BooleanCriteria { int ordinal; std::function<bool(bool)> booleanPredicate; };
using Criteria = std::variant<IntCriteria, StringCriteria, BooleanCriteria>;

class Query {
public:
    class Node {
    public:
        Node& createChild(int ordinal) {
            if (children.count(ordinal)) throw std::invalid_argument("A child already exists at ordinal " + std::to_string(ordinal));
            return *(children[ordinal] = std::make_unique<Node>());
        }
        Node& addCriteria(Criteria c) { criteria.push_back(std::move(c)); return *this; }
        std::map<int, std::unique_ptr<Node>> children;
        std::vector<Criteria> criteria;
    };
    explicit Query(std::string name) : tableName(std::move(name)) {}
    std::string tableName;
    Node rootNode;
};

struct Success { std::shared_ptr<InMemoryTable> resultSet; };
struct Failure { std::string message; };
using QueryResult = std::variant<Success, Failure>;

// physical layout of the registered tables on the engine side (include/colq.h "Host-resident columns" and
// "Dictionary-encoded StringColumn"); results are identical in every combination
enum class Residency { Device, Host };
struct Options {
    Residency residency = Residency::Device;  // Host: columns stay in pinned off-heap buffers, streamed over PCIe, promoted on first scan
    bool dictionary = false;                  // string columns as int32 codes + distinct values; enables StringLambda criteria
};

// ---- E/DataSystemSerialIndices.java:14-102 over libcolq.so
class DataSystemColq {
public:
    explicit DataSystemColq(int device = 0, Options options = {}) : options_(options) {
        if (colq_create(device, &ctx_) != COLQ_OK)
            throw std::runtime_error("colq_create failed: no usable sm_100 GPU (libcolq has no CPU fallback)");
    }
    ~DataSystemColq() {
        for (void* p : hostBuffers_) colq_host_free(ctx_, p);
        colq_destroy(ctx_);
    }
    DataSystemColq(const DataSystemColq&) = delete;

    void registerTable(const std::string& name, std::shared_ptr<InMemoryTable> table) { tables_[name] = std::move(table); }

    QueryResult execute(const Query& query) {
        auto it = tables_.find(query.tableName);
        if (it == tables_.end())
            return Failure{"The query targets the table '" + query.tableName + "' but that table is not registered"};
        InMemoryTable& table = *it->second;
        syncTables();
        colq_query* q = nullptr;
        check(colq_query_create(ctx_, query.tableName.c_str(), &q));
        struct Guard { colq_query* q; ~Guard() { colq_query_destroy(q); } } guard{q};
        if (std::string why = translate(q, query.rootNode, 0, &table); !why.empty()) return Failure{why};
        BitSet bits;
        bits.words.assign((size_t)((table.size() + 63) / 64), 0);
        int64_t count = 0;
        colq_status st = colq_execute(ctx_, q, bits.words.data(), (int64_t)bits.words.size(), nullptr, 0, &count, nullptr);
        if (st == COLQ_FAILURE) return Failure{colq_last_error(ctx_)};
        check(st);
        return Success{table.subset(bits)};  // table.subset(matchingRows) (:100)
    }

private:
    void check(colq_status st) {
        switch (st) {
            case COLQ_OK: return;
            case COLQ_THROW_INDEX_OOB: throw std::out_of_range(std::string("IndexOutOfBoundsException: ") + colq_last_error(ctx_));
            case COLQ_THROW_ILLEGAL_ARG: throw std::invalid_argument(colq_last_error(ctx_));
            default: throw std::runtime_error(std::string("libcolq status ") + std::to_string((int)st) + ": " + colq_last_error(ctx_));
        }
    }
    // returns a Failure message for an opaque predicate the engine cannot run, else ""
    std::string translate(colq_query* q, const Query::Node& node, int id, InMemoryTable* table) {
        for (const Criteria& c : node.criteria) {
            if (auto* ic = std::get_if<IntCriteria>(&c)) {
                check(colq_query_criteria_i32_range(q, id, ic->ordinal, ic->integerPredicate.lo, ic->integerPredicate.hi));
                continue;
            }
            if (auto* bc = std::get_if<BooleanCriteria>(&c)) {
                check(colq_query_criteria_bool(q, id, bc->ordinal, bc->booleanPredicate(false) ? 1 : 0, bc->booleanPredicate(true) ? 1 : 0));
                continue;
            }
            const auto& sc = std::get<StringCriteria>(c);
            if (auto* op = std::get_if<StringOp>(&sc.stringPredicate)) {
                check(colq_query_criteria_str(q, id, sc.ordinal, op->op, reinterpret_cast<const uint8_t*>(op->value.data()),
                                              (int32_t)op->value.size()));
                continue;
            }
            auto dv = table ? dictValues_.find({table, sc.ordinal}) : dictValues_.end();
            if (dv == dictValues_.end())
                return "The criterion on ordinal " + std::to_string(sc.ordinal) + " is an opaque string predicate; the GPU engine runs those only over "
                       "dictionary-encoded columns (Options::dictionary) and has no CPU fallback.";
            const auto& fn = std::get<StringLambda>(sc.stringPredicate);
            std::vector<uint64_t> words(dv->second.size() / 64 + 1, 0);
            for (size_t d = 0; d < dv->second.size(); ++d)
                if (fn(dv->second[d])) words[d >> 6] |= uint64_t(1) << (d & 63);
            check(colq_query_criteria_str_accept(q, id, sc.ordinal, words.data(), (int64_t)dv->second.size()));
        }
        for (const auto& [ordinal, child] : node.children) {
            int cid = 0;
            check(colq_query_child(q, id, ordinal, &cid));
            InMemoryTable* childTable = nullptr;
            if (table && ordinal >= 0 && ordinal < table->width())
                if (auto* a = std::get_if<AssociationColumn>(&table->columns[(size_t)ordinal])) childTable = a->associatedEntity;
            if (std::string why = translate(q, *child, cid, childTable); !why.empty()) return why;
        }
        return {};
    }
    // a pinned, device-mapped copy of `n` elements, padded for whole-line reads (freed with the data system)
    template <class T>
    T* hostCopy(const T* src, size_t n, int64_t* capacityBytes) {
        const size_t bytes = (n * sizeof(T) + 15) / 16 * 16 + 64;
        void* p = nullptr;
        check(colq_host_alloc(ctx_, (int64_t)bytes, &p));
        hostBuffers_.push_back(p);
        std::memset(p, 0, bytes);
        if (n) std::memcpy(p, src, n * sizeof(T));
        *capacityBytes = (int64_t)bytes;
        return static_cast<T*>(p);
    }
    void syncTables() {
        std::vector<InMemoryTable*> todo, seen;
        for (auto& kv : tables_) todo.push_back(kv.second.get());
        while (!todo.empty()) {
            InMemoryTable* t = todo.back(); todo.pop_back();
            bool dup = false;
            for (auto* s : seen) dup |= (s == t);
            if (dup) continue;
            seen.push_back(t);
            for (auto& c : t->columns) if (auto* a = std::get_if<AssociationColumn>(&c)) todo.push_back(a->associatedEntity);
        }
        for (auto* t : seen) if (!handles_.count(t)) {
            colq_table h;
            check(colq_table_create(ctx_, t->size(), COLQ_REPLICATED, 0, &h));
            handles_[t] = h; uploaded_[t] = 0;
        }
        for (auto* t : seen) for (int o = uploaded_[t]; o < t->width(); ++o) {
            const colq_table h = handles_[t];
            const bool host = options_.residency == Residency::Host && t->size() > 0;
            int64_t cap = 0, cap2 = 0;
            if (auto* c = std::get_if<IntegerColumn>(&t->columns[(size_t)o])) {
                if (host) {
                    const int32_t* pinned = hostCopy(c->ints.data(), c->ints.size(), &cap);  // (sets cap: keep it a separate statement)
                    check(colq_col_i32_host(ctx_, h, o, pinned, cap, (int64_t)c->ints.size()));
                } else check(colq_col_i32(ctx_, h, o, c->ints.data(), (int64_t)c->ints.size()));
            } else if (auto* s = std::get_if<StringColumn>(&t->columns[(size_t)o])) {
                const int64_t n = (int64_t)s->strings.size();
                if (options_.dictionary) {
                    // dictionary-encode: distinct values in first-appearance order
                    std::map<std::string, int32_t> index;
                    std::vector<std::string>& values = dictValues_[{t, o}];
                    std::vector<int32_t> codes;
                    for (const std::string& v : s->strings) {
                        auto [it, fresh] = index.emplace(v, (int32_t)values.size());
                        if (fresh) values.push_back(v);
                        codes.push_back(it->second);
                    }
                    std::vector<uint32_t> off(values.size() + 1, 0);
                    std::string bytes;
                    for (size_t i = 0; i < values.size(); ++i) { bytes += values[i]; off[i + 1] = (uint32_t)bytes.size(); }
                    const auto* db = reinterpret_cast<const uint8_t*>(bytes.data());
                    if (host) {
                        const int32_t* pinned = hostCopy(codes.data(), codes.size(), &cap);
                        check(colq_col_str_dict_host(ctx_, h, o, pinned, cap, n, off.data(), db, (int64_t)values.size(), (int64_t)bytes.size()));
                    } else check(colq_col_str_dict(ctx_, h, o, codes.data(), n, off.data(), db, (int64_t)values.size(), (int64_t)bytes.size()));
                    continue;
                }
                std::vector<uint32_t> off(s->strings.size() + 1, 0);
                std::string bytes;
                for (size_t i = 0; i < s->strings.size(); ++i) { bytes += s->strings[i]; off[i + 1] = (uint32_t)bytes.size(); }
                const auto* sb = reinterpret_cast<const uint8_t*>(bytes.data());
                if (host) {
                    uint32_t* ho = hostCopy(off.data(), off.size(), &cap);
                    uint8_t* hb = hostCopy(sb, bytes.size(), &cap2);
                    check(colq_col_str_host(ctx_, h, o, ho, cap, hb, cap2, n, (int64_t)bytes.size()));
                } else check(colq_col_str(ctx_, h, o, off.data(), sb, n, (int64_t)bytes.size()));
            } else if (auto* b = std::get_if<BooleanColumn>(&t->columns[(size_t)o])) check(colq_col_bool(ctx_, h, o, b->bools.data(), (int64_t)b->bools.size()));
        }
        for (auto* t : seen) for (int o = uploaded_[t]; o < t->width(); ++o) {
            auto* a = std::get_if<AssociationColumn>(&t->columns[(size_t)o]);
            if (!a || !a->forward) continue;
            bool toOne = true;
            for (auto& as : a->associations) toOne &= !std::holds_alternative<Many>(as);
            const colq_table hx = handles_[t], hy = handles_[a->associatedEntity];
            if (toOne) {
                std::vector<int32_t> fk;
                for (auto& as : a->associations) fk.push_back(std::holds_alternative<One>(as) ? std::get<One>(as).idx : -1);
                int64_t cap = 0;
                if (options_.residency == Residency::Host && !fk.empty()) {
                    const int32_t* pinned = hostCopy(fk.data(), fk.size(), &cap);
                    check(colq_associate_fk_host(ctx_, hx, o, hy, a->reverseOrdinal, pinned, cap, (int64_t)fk.size()));
                } else check(colq_associate_fk(ctx_, hx, o, hy, a->reverseOrdinal, fk.data(), (int64_t)fk.size()));
            } else {
                std::vector<int64_t> off{0};
                std::vector<int32_t> tgt;
                for (auto& as : a->associations) {
                    if (auto* one = std::get_if<One>(&as)) tgt.push_back(one->idx);
                    else if (auto* m = std::get_if<Many>(&as)) tgt.insert(tgt.end(), m->indices.begin(), m->indices.end());
                    off.push_back((int64_t)tgt.size());
                }
                check(colq_associate_csr(ctx_, hx, o, hy, a->reverseOrdinal, off.data(), tgt.data(), (int64_t)a->associations.size(), (int64_t)tgt.size()));
            }
        }
        for (auto* t : seen) uploaded_[t] = t->width();
        for (auto& kv : tables_) {
            const colq_table h = handles_[kv.second.get()];
            if (!registered_.count(kv.first) || registered_[kv.first] != h) { check(colq_register(ctx_, kv.first.c_str(), h)); registered_[kv.first] = h; }
        }
    }

    colq_ctx* ctx_ = nullptr;
    Options options_;
    std::vector<void*> hostBuffers_;
    std::map<std::pair<InMemoryTable*, int>, std::vector<std::string>> dictValues_;
    std::map<std::string, std::shared_ptr<InMemoryTable>> tables_;
    std::map<InMemoryTable*, colq_table> handles_;
    std::map<InMemoryTable*, int> uploaded_;
    std::map<std::string, colq_table> registered_;
};

}  // namespace colq
