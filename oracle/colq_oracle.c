/*
 * colq_oracle.c -- literal CPU restatement of the reference engine.  See colq_oracle.h for the
 * "test infrastructure only" contract and the parity-pinning status.
 *
 * The restatement is deliberately literal: per-node BitSet of uint64 words, self filter by an
 * AND of per-row predicates, upward pruning by PUSHING the child's matching rows through the
 * reverse association column and AND-ing into the parent (never the GPU's fused pull form), then an
 * ascending scan for the subset.  That makes it an independent check on the CUDA path.
 *
 * Citation shorthands (relative to /root/reference/):
 *   E  = data-system-serial-indices-arrays/src/main/java/dgroomes/data_system_serial_indices_arrays
 *   M  = data-model-in-memory/src/main/java/dgroomes/in_memory
 *   DS = data-system/src/main/java/dgroomes/data_system
 */
#include "colq_oracle.h"

#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#ifdef _OPENMP
#include <omp.h>
#endif

/* ------------------------------------------------------------------ data model (M/) */

enum { COL_INT = 0, COL_STR = 1, COL_BOOL = 2, COL_ASSOC = 3 };

typedef struct {
    int kind;
    int64_t n;            /* height() */
    int32_t *ints;        /* IntegerColumn.ints   M/InMemoryColumn.java:46 */
    uint32_t *soff;       /* StringColumn.strings M/InMemoryColumn.java:64, as offsets+UTF-8 bytes */
    uint8_t *sbytes;
    uint8_t *bools;       /* BooleanColumn.bools  M/InMemoryColumn.java:28 */
    /* AssociationColumn  M/InMemoryColumn.java:85-138 */
    int assoc_table;      /* associatedEntity */
    uint8_t *akind;       /* per row: 0 None, 1 One, 2 Many  (DS/Association.java:27-51) */
    int64_t *aoff;
    int32_t *atgt;
    int rev_table, rev_ordinal; /* reverseAssociatedColumn (set by associateTo, M/InMemoryTable.java:83-85) */
} column;

typedef struct {
    column *cols;
    int ncols;
} table;

typedef struct {
    char *name;
    int table;
} registration;

/* ------------------------------------------------------------------ query (DS/Query.java) */

typedef struct {
    int ordinal;
    int is_str;           /* Criteria.StringCriteria vs Criteria.IntCriteria (DS/Criteria.java:17,19) */
    int32_t lo, hi;
    int op;
    uint8_t *needle;
    int32_t nlen;
    /* 8(f4) extension: a Predicate<Boolean> as its truth table -- the where() the reference declares
       (DS/ColumnFilterable.java:20-22) and its Verifier refuses (E/Verifier.java:82-84) */
    int is_bool, accept_false, accept_true;
} criterion;

typedef struct {
    criterion *crit;
    int ncrit;
    int *child_ordinal;   /* childrenByOrdinal (DS/Query.java:27) */
    int *child_node;
    int nchild;
} qnode;

struct orc_query {
    char *table_name;
    qnode *nodes;
    int nnodes;
};

/* ------------------------------------------------------------------ execution context (E/ExecutionContext.java) */

typedef struct {
    int table;
    int parent;                  /* index into exec nodes, -1 for the root */
    int up_table, up_ordinal;    /* associationToParent: a column of THIS node's table (:65) */
    criterion **preds;           /* columnPredicates, bound to columns (:39,60-62) */
    int npreds;
    uint64_t *bits;              /* matchingBits (:41) */
    int64_t size;
    int *children;
    int nchildren;
} xnode;

struct orc_system {
    table *tables;
    int ntables;
    registration *regs;
    int nregs;
    char msg[512];
    int64_t last_card[64];
    int last_nnodes;
};

/* ------------------------------------------------------------------ small helpers */

static void *xmalloc(size_t n) {
    void *p = malloc(n ? n : 1);
    if (!p) { fprintf(stderr, "oracle: out of memory\n"); abort(); }
    return p;
}
static void *xcalloc(size_t n, size_t m) {
    void *p = calloc(n ? n : 1, m ? m : 1);
    if (!p) { fprintf(stderr, "oracle: out of memory\n"); abort(); }
    return p;
}
static void *xdup(const void *src, size_t n) {
    void *p = xmalloc(n);
    if (n) memcpy(p, src, n);
    return p;
}

/* java.util.BitSet word layout: bit i lives in words[i >> 6] at (1L << (i & 63)) */
static inline void bs_set(uint64_t *w, int64_t i) { w[i >> 6] |= (uint64_t)1 << (i & 63); }
static inline int bs_get(const uint64_t *w, int64_t i) { return (int)((w[i >> 6] >> (i & 63)) & 1u); }
static inline int64_t bs_nwords(int64_t nbits) { return (nbits + 63) >> 6; }

orc_system *orc_system_new(void) { return (orc_system *)xcalloc(1, sizeof(orc_system)); }

static void column_free(column *c) {
    free(c->ints); free(c->soff); free(c->sbytes); free(c->bools);
    free(c->akind); free(c->aoff); free(c->atgt);
}

void orc_system_free(orc_system *s) {
    if (!s) return;
    for (int t = 0; t < s->ntables; t++) {
        for (int c = 0; c < s->tables[t].ncols; c++) column_free(&s->tables[t].cols[c]);
        free(s->tables[t].cols);
    }
    free(s->tables);
    for (int r = 0; r < s->nregs; r++) free(s->regs[r].name);
    free(s->regs);
    free(s);
}

const char *orc_last_message(const orc_system *s) { return s->msg; }
void orc_free(void *p) { free(p); }

int orc_table_new(orc_system *s) {
    s->tables = (table *)realloc(s->tables, sizeof(table) * (size_t)(s->ntables + 1));
    memset(&s->tables[s->ntables], 0, sizeof(table));
    return s->ntables++;
}

static column *push_column(orc_system *s, int t) {
    table *tb = &s->tables[t];
    tb->cols = (column *)realloc(tb->cols, sizeof(column) * (size_t)(tb->ncols + 1));
    column *c = &tb->cols[tb->ncols++];
    memset(c, 0, sizeof(*c));
    c->rev_table = c->rev_ordinal = -1;
    return c;
}

int orc_table_add_ints(orc_system *s, int t, const int32_t *v, int64_t n) {
    column *c = push_column(s, t);
    c->kind = COL_INT; c->n = n;
    c->ints = (int32_t *)xdup(v, sizeof(int32_t) * (size_t)n);
    return s->tables[t].ncols - 1;
}

int orc_table_add_strings(orc_system *s, int t, const uint32_t *off, const uint8_t *bytes, int64_t n) {
    column *c = push_column(s, t);
    c->kind = COL_STR; c->n = n;
    c->soff = (uint32_t *)xdup(off, sizeof(uint32_t) * (size_t)(n + 1));
    c->sbytes = (uint8_t *)xdup(bytes, off[n]);
    return s->tables[t].ncols - 1;
}

int orc_table_add_bools(orc_system *s, int t, const uint8_t *v, int64_t n) {
    column *c = push_column(s, t);
    c->kind = COL_BOOL; c->n = n;
    c->bools = (uint8_t *)xdup(v, (size_t)n);
    return s->tables[t].ncols - 1;
}

/* InMemoryTable.size(): length of column 0 (M/InMemoryTable.java:92-101) */
int64_t orc_table_size(const orc_system *s, int t) {
    const table *tb = &s->tables[t];
    return tb->ncols ? tb->cols[0].n : -1;
}
int orc_table_width(const orc_system *s, int t) { return s->tables[t].ncols; }

/*
 * InMemoryTable.associateTo (M/InMemoryTable.java:44-90): append the forward column to X (:48),
 * build y -> [x...] with x ascending (:61-73), classify each y as None/One/Many by list length
 * (:75-82), cross-link both columns (:83-85), append the reverse column to Y (:88).
 */
int orc_table_associate(orc_system *s, int x, int y, const uint8_t *kind, const int64_t *off,
                        const int32_t *tgt, int64_t n, int *x_ord, int *y_ord) {
    int64_t ysize = orc_table_size(s, y);
    int64_t nnz = off[n];
    /* a target outside [0, ysize) makes yIndexToXAssociations.get(yIndex) null -> NPE on .add (:70-71) */
    for (int64_t i = 0; i < n; i++) {
        if (kind[i] == 0) continue;  /* Association.None contributes `new int[]{}` (:65) */
        for (int64_t e = off[i]; e < off[i + 1]; e++)
            if (tgt[e] < 0 || tgt[e] >= ysize) {
                snprintf(s->msg, sizeof s->msg, "NullPointerException: association target %d outside the associated table (size %lld)",
                         tgt[e], (long long)ysize);
                return ORC_THROW_NULL;
            }
    }
    column *f = push_column(s, x);
    int xo = s->tables[x].ncols - 1;
    f->kind = COL_ASSOC; f->n = n; f->assoc_table = y;
    f->akind = (uint8_t *)xdup(kind, (size_t)n);
    f->aoff = (int64_t *)xdup(off, sizeof(int64_t) * (size_t)(n + 1));
    f->atgt = (int32_t *)xdup(tgt, sizeof(int32_t) * (size_t)nnz);

    /* transpose: counting pass then ascending-x fill */
    int64_t *roff = (int64_t *)xcalloc((size_t)ysize + 1, sizeof(int64_t));
    for (int64_t i = 0; i < n; i++) {
        if (kind[i] == 0) continue;
        for (int64_t e = off[i]; e < off[i + 1]; e++) roff[tgt[e] + 1]++;
    }
    for (int64_t j = 0; j < ysize; j++) roff[j + 1] += roff[j];
    int64_t rnnz = roff[ysize];
    int32_t *rtgt = (int32_t *)xmalloc(sizeof(int32_t) * (size_t)rnnz);
    int64_t *cur = (int64_t *)xdup(roff, sizeof(int64_t) * (size_t)(ysize + 1));
    for (int64_t i = 0; i < n; i++) {
        if (kind[i] == 0) continue;
        for (int64_t e = off[i]; e < off[i + 1]; e++) rtgt[cur[tgt[e]]++] = (int32_t)i;
    }
    free(cur);
    uint8_t *rkind = (uint8_t *)xmalloc((size_t)ysize);
    for (int64_t j = 0; j < ysize; j++) {
        int64_t cnt = roff[j + 1] - roff[j];
        rkind[j] = cnt == 0 ? 0 : (cnt == 1 ? 1 : 2);
    }
    /* push_column may realloc the X table's column array when x == y, so re-fetch `f` afterwards */
    column *r = push_column(s, y);
    int yo = s->tables[y].ncols - 1;
    r->kind = COL_ASSOC; r->n = ysize; r->assoc_table = x;
    r->akind = rkind; r->aoff = roff; r->atgt = rtgt;
    r->rev_table = x; r->rev_ordinal = xo;
    f = &s->tables[x].cols[xo];
    f->rev_table = y; f->rev_ordinal = yo;
    if (x_ord) *x_ord = xo;
    if (y_ord) *y_ord = yo;
    return ORC_SUCCESS;
}

/* DataSystemSerialIndices.register: HashMap.put, later put wins (E/DataSystemSerialIndices.java:27-29) */
void orc_register(orc_system *s, const char *name, int t) {
    for (int r = 0; r < s->nregs; r++)
        if (strcmp(s->regs[r].name, name) == 0) { s->regs[r].table = t; return; }
    s->regs = (registration *)realloc(s->regs, sizeof(registration) * (size_t)(s->nregs + 1));
    s->regs[s->nregs].name = strdup(name);
    s->regs[s->nregs].table = t;
    s->nregs++;
}

/* ------------------------------------------------------------------ query building */

orc_query *orc_query_new(const char *table_name) {
    orc_query *q = (orc_query *)xcalloc(1, sizeof(*q));
    q->table_name = strdup(table_name);
    q->nodes = (qnode *)xcalloc(1, sizeof(qnode));
    q->nnodes = 1;  /* rootNode (DS/Query.java:22-25) */
    return q;
}

void orc_query_free(orc_query *q) {
    if (!q) return;
    for (int i = 0; i < q->nnodes; i++) {
        for (int c = 0; c < q->nodes[i].ncrit; c++) free(q->nodes[i].crit[c].needle);
        free(q->nodes[i].crit); free(q->nodes[i].child_ordinal); free(q->nodes[i].child_node);
    }
    free(q->nodes); free(q->table_name); free(q);
}

/* Query.Node.createChild: duplicate ordinal -> IllegalArgumentException (DS/Query.java:31-38) */
int orc_query_create_child(orc_query *q, int parent, int ordinal) {
    qnode *p = &q->nodes[parent];
    for (int i = 0; i < p->nchild; i++)
        if (p->child_ordinal[i] == ordinal) return -ORC_THROW_ILLEGAL_ARG;
    q->nodes = (qnode *)realloc(q->nodes, sizeof(qnode) * (size_t)(q->nnodes + 1));
    memset(&q->nodes[q->nnodes], 0, sizeof(qnode));
    p = &q->nodes[parent];
    p->child_ordinal = (int *)realloc(p->child_ordinal, sizeof(int) * (size_t)(p->nchild + 1));
    p->child_node = (int *)realloc(p->child_node, sizeof(int) * (size_t)(p->nchild + 1));
    p->child_ordinal[p->nchild] = ordinal;
    p->child_node[p->nchild] = q->nnodes;
    p->nchild++;
    return q->nnodes++;
}

static criterion *push_criterion(orc_query *q, int node) {
    qnode *n = &q->nodes[node];
    n->crit = (criterion *)realloc(n->crit, sizeof(criterion) * (size_t)(n->ncrit + 1));
    criterion *c = &n->crit[n->ncrit++];
    memset(c, 0, sizeof(*c));
    return c;
}

void orc_query_add_int_range(orc_query *q, int node, int ordinal, int32_t lo, int32_t hi) {
    criterion *c = push_criterion(q, node);
    c->ordinal = ordinal; c->is_str = 0; c->lo = lo; c->hi = hi;
}

void orc_query_add_str(orc_query *q, int node, int ordinal, int op, const uint8_t *needle, int32_t len) {
    criterion *c = push_criterion(q, node);
    c->ordinal = ordinal; c->is_str = 1; c->op = op; c->nlen = len;
    c->needle = (uint8_t *)xdup(needle, (size_t)len);
}

void orc_query_add_bool(orc_query *q, int node, int ordinal, int accept_false, int accept_true) {
    criterion *c = push_criterion(q, node);
    c->ordinal = ordinal; c->is_bool = 1; c->accept_false = accept_false != 0; c->accept_true = accept_true != 0;
}

/* ------------------------------------------------------------------ predicates */

/*
 * java.lang.String.compareTo compares UTF-16 code units (JDK String.compareTo; called by the reference's lambdas at
 * QueryTest.java:124-125).  The oracle does it literally: decode one UTF-8 code point from each side (the columns
 * hold well-formed UTF-8, the encoding the Java shim writes), expand it to its one or two UTF-16 code units and compare
 * unit by unit -- deliberately NOT the lead-byte key trick the GPU kernel uses, so the two are independent.
 */
static int utf8_next(const uint8_t *s, int64_t len, int64_t *i, uint16_t out[2]) {
    uint32_t b = s[*i], cp;
    int extra = b < 0x80 ? 0 : (b < 0xE0 ? 1 : (b < 0xF0 ? 2 : 3));
    cp = extra == 0 ? b : (extra == 1 ? (b & 0x1F) : (extra == 2 ? (b & 0x0F) : (b & 0x07)));
    for (int k = 1; k <= extra; k++) cp = (cp << 6) | ((*i + k < len ? s[*i + k] : 0) & 0x3F);
    *i += extra + 1;
    if (cp >= 0x10000) {
        cp -= 0x10000;
        out[0] = (uint16_t)(0xD800 + (cp >> 10));
        out[1] = (uint16_t)(0xDC00 + (cp & 0x3FF));
        return 2;
    }
    out[0] = (uint16_t)cp;
    return 1;
}

static int java_compare_to(const uint8_t *a, int64_t la, const uint8_t *b, int64_t lb, int *sign_only) {
    (void)sign_only;
    int64_t i = 0, j = 0;
    uint16_t ua[2], ub[2];
    int na = 0, nb = 0, pa = 0, pb = 0;  /* pending code units of the current code point on each side */
    for (;;) {
        if (pa == na) { if (i >= la) { na = 0; } else { na = utf8_next(a, la, &i, ua); } pa = 0; }
        if (pb == nb) { if (j >= lb) { nb = 0; } else { nb = utf8_next(b, lb, &j, ub); } pb = 0; }
        if (na == 0 || nb == 0) return (na == 0 && nb == 0) ? 0 : (na == 0 ? -1 : 1);  /* length difference */
        if (ua[pa] != ub[pb]) return ua[pa] < ub[pb] ? -1 : 1;
        pa++; pb++;
    }
}

static int bytes_contains(const uint8_t *h, int64_t hl, const uint8_t *n, int64_t nl) {
    if (nl == 0) return 1;  /* "".contains("") and s.contains("") are true */
    if (nl > hl) return 0;
    for (int64_t p = 0; p + nl <= hl; p++)
        if (h[p] == n[0] && memcmp(h + p, n, (size_t)nl) == 0) return 1;
    return 0;
}

static inline int str_test(const criterion *c, const uint8_t *s, int64_t len) {
    switch (c->op) {
        case ORC_STR_EQ: return len == c->nlen && memcmp(s, c->needle, (size_t)len) == 0;
        case ORC_STR_NE: return !(len == c->nlen && memcmp(s, c->needle, (size_t)len) == 0);
        case ORC_STR_CONTAINS: return bytes_contains(s, len, c->needle, c->nlen);
        case ORC_STR_CMP_GT: return java_compare_to(s, len, c->needle, c->nlen, 0) > 0;
        case ORC_STR_CMP_LT: return java_compare_to(s, len, c->needle, c->nlen, 0) < 0;
        case ORC_STR_CMP_GE: return java_compare_to(s, len, c->needle, c->nlen, 0) >= 0;
        case ORC_STR_CMP_LE: return java_compare_to(s, len, c->needle, c->nlen, 0) <= 0;
        case ORC_STR_STARTS_WITH: return len >= c->nlen && memcmp(s, c->needle, (size_t)c->nlen) == 0;
        case ORC_STR_ENDS_WITH: return len >= c->nlen && memcmp(s + len - c->nlen, c->needle, (size_t)c->nlen) == 0;
    }
    return 0;
}

/* idx -> predicate.test(ints[idx]) / predicate.test(strings[idx])  (M/InMemoryColumn.java:53-56,71-74) */
static inline int pred_test(const orc_system *s, const xnode *n, const criterion *c, int64_t i) {
    const column *col = &s->tables[n->table].cols[c->ordinal];
    if (c->is_bool) return col->bools[i] ? c->accept_true : c->accept_false;
    if (c->is_str) {
        uint32_t a = col->soff[i], b = col->soff[i + 1];
        return str_test(c, col->sbytes + a, (int64_t)b - a);
    }
    int32_t v = col->ints[i];
    return v >= c->lo && v <= c->hi;
}

/* ------------------------------------------------------------------ verify (E/Verifier.java:40-111) */

static const char *java_class_name(int kind) {
    switch (kind) {
        case COL_INT: return "dgroomes.in_memory.InMemoryColumn$IntegerColumn";
        case COL_STR: return "dgroomes.in_memory.InMemoryColumn$StringColumn";
        case COL_BOOL: return "dgroomes.in_memory.InMemoryColumn$BooleanColumn";
        default: return "dgroomes.in_memory.InMemoryColumn$AssociationColumn";
    }
}

typedef struct {
    xnode *nodes;
    int n, cap;
} xctx;

static int xctx_new_node(orc_system *s, xctx *x, int t, int parent, int up_table, int up_ordinal) {
    if (x->n == x->cap) {
        x->cap = x->cap ? x->cap * 2 : 8;
        x->nodes = (xnode *)realloc(x->nodes, sizeof(xnode) * (size_t)x->cap);
    }
    xnode *n = &x->nodes[x->n];
    memset(n, 0, sizeof(*n));
    n->table = t; n->parent = parent; n->up_table = up_table; n->up_ordinal = up_ordinal;
    n->size = orc_table_size(s, t);                                  /* new BitSet(table.size()) (:57) */
    n->bits = (uint64_t *)xcalloc((size_t)bs_nwords(n->size > 0 ? n->size : 0) + 1, sizeof(uint64_t));
    return x->n++;
}

static void xctx_free(xctx *x) {
    for (int i = 0; i < x->n; i++) { free(x->nodes[i].bits); free(x->nodes[i].preds); free(x->nodes[i].children); }
    free(x->nodes);
}

static int verify(orc_system *s, const orc_query *q, int root_table, xctx *x) {
    /* record NodeNode(queryNode, executionNode); toVisit.add(...) = tail, toVisit.pop() = head -> BFS (:49-54) */
    int *fifo_q = (int *)xmalloc(sizeof(int) * (size_t)q->nnodes);
    int *fifo_x = (int *)xmalloc(sizeof(int) * (size_t)q->nnodes);
    int head = 0, tail = 0, rc = ORC_SUCCESS;
    fifo_q[tail] = 0; fifo_x[tail] = xctx_new_node(s, x, root_table, -1, -1, -1); tail++;
    while (head < tail) {
        const qnode *qn = &q->nodes[fifo_q[head]];
        int xi = fifo_x[head]; head++;
        int t = x->nodes[xi].table;
        const table *tb = &s->tables[t];
        for (int k = 0; k < qn->ncrit; k++) {
            criterion *c = &qn->crit[k];
            if (tb->ncols < c->ordinal) {  /* sic: `<`, so ordinal == width falls through to get() (:62) */
                snprintf(s->msg, sizeof s->msg, "The query ordinal '%d' is out of bounds for the table with %d columns",
                         c->ordinal, tb->ncols);
                rc = ORC_FAILURE; goto done;
            }
            if (c->ordinal < 0 || c->ordinal >= tb->ncols) {  /* columns().get(ordinal) (:67) */
                snprintf(s->msg, sizeof s->msg, "IndexOutOfBoundsException: Index %d out of bounds for length %d",
                         c->ordinal, tb->ncols);
                rc = ORC_THROW_INDEX_OOB; goto done;
            }
            const column *col = &tb->cols[c->ordinal];
            switch (col->kind) {   /* switch (column.filterableType()) (:71-90) */
                case COL_STR:
                    if (!c->is_str || c->is_bool) {
                        snprintf(s->msg, sizeof s->msg, "The column is a string column but the criterion is not a string predicate.");
                        rc = ORC_FAILURE; goto done;
                    }
                    break;
                case COL_INT:
                    if (c->is_str || c->is_bool) {
                        snprintf(s->msg, sizeof s->msg, "The column is an integer column but the criterion is not an integer predicate.");
                        rc = ORC_FAILURE; goto done;
                    }
                    break;
                case COL_BOOL:
                    if (c->is_bool) break;   /* the extension; every criterion the reference can express fails (:82-84) */
                    snprintf(s->msg, sizeof s->msg, "Boolean columns are not supported yet.");
                    rc = ORC_FAILURE; goto done;
                default:
                    snprintf(s->msg, sizeof s->msg, "Association columns can't be matched on with a scalar criteria.");
                    rc = ORC_FAILURE; goto done;
            }
            xnode *xn = &x->nodes[xi];
            xn->preds = (criterion **)realloc(xn->preds, sizeof(criterion *) * (size_t)(xn->npreds + 1));
            xn->preds[xn->npreds++] = c;   /* addColumnPredicate (:92) */
        }
        for (int k = 0; k < qn->nchild; k++) {   /* Map.copyOf iteration: order unspecified, result order-free (:94-107) */
            int ordinal = qn->child_ordinal[k];
            if (ordinal < 0 || ordinal >= tb->ncols) {  /* columns().get(ordinal) unchecked (:100) */
                snprintf(s->msg, sizeof s->msg, "IndexOutOfBoundsException: Index %d out of bounds for length %d",
                         ordinal, tb->ncols);
                rc = ORC_THROW_INDEX_OOB; goto done;
            }
            const column *col = &tb->cols[ordinal];
            if (col->kind != COL_ASSOC) {
                snprintf(s->msg, sizeof s->msg, "The column at ordinal %d is not an association column. It is a %s",
                         ordinal, java_class_name(col->kind));
                rc = ORC_FAILURE; goto done;
            }
            if (col->rev_table < 0) {  /* reverseAssociatedColumn() never set (M/InMemoryColumn.java:122-126) */
                snprintf(s->msg, sizeof s->msg, "IllegalStateException: reverseAssociatedColumn was never set");
                rc = ORC_THROW_ILLEGAL_STATE; goto done;
            }
            /* createChildNode: Node(assoc.associatedEntity(), this, assoc.reverseAssociatedColumn()) (E/ExecutionContext.java:64-68) */
            int child = xctx_new_node(s, x, col->assoc_table, xi, col->rev_table, col->rev_ordinal);
            xnode *xn = &x->nodes[xi];
            xn->children = (int *)realloc(xn->children, sizeof(int) * (size_t)(xn->nchildren + 1));
            xn->children[xn->nchildren++] = child;
            fifo_q[tail] = qn->child_node[k]; fifo_x[tail] = child; tail++;
        }
    }
done:
    free(fifo_q); free(fifo_x);
    return rc;
}

/* ------------------------------------------------------------------ filterSelf / filterParent */

/* ExecutionContext.Node.filterSelf (E/ExecutionContext.java:79-94) */
static void filter_self(const orc_system *s, xnode *n, int nthreads) {
    int64_t size = n->size;
    if (n->npreds == 0) {                       /* matchingBits.set(0, table.size()) (:83-87) */
        for (int64_t w = 0; w < (size >> 6); w++) n->bits[w] = ~(uint64_t)0;
        if (size & 63) n->bits[size >> 6] = (((uint64_t)1 << (size & 63)) - 1);
        return;
    }
    /* for i in [0,size): if (combinedPredicate.test(i)) matchingBits.set(i)  (:91-93);
       IntPredicate.and short-circuits left to right (:81). Threads own whole 64-row words. */
    int64_t nwords = bs_nwords(size);
    (void)nthreads;
#ifdef _OPENMP
#pragma omp parallel for schedule(static) num_threads(nthreads) if (nthreads > 1)
#endif
    for (int64_t w = 0; w < nwords; w++) {
        int64_t lo = w << 6, hi = lo + 64 < size ? lo + 64 : size;
        uint64_t word = 0;
        for (int64_t i = lo; i < hi; i++) {
            int ok = 1;
            for (int k = 0; k < n->npreds && ok; k++) ok = pred_test(s, n, n->preds[k], i);
            if (ok) word |= (uint64_t)1 << (i & 63);
        }
        n->bits[w] = word;
    }
}

/* ExecutionContext.Node.filterParent (E/ExecutionContext.java:100-122) */
static int filter_parent(const orc_system *s, xctx *x, int ni, int nthreads) {
    xnode *n = &x->nodes[ni];
    if (n->parent < 0) return ORC_SUCCESS;      /* the root has no parent (:101) */
    xnode *p = &x->nodes[n->parent];
    const column *up = &s->tables[n->up_table].cols[n->up_ordinal];
    int64_t pw = bs_nwords(p->size);
    uint64_t *reach = (uint64_t *)xcalloc((size_t)pw + 1, sizeof(uint64_t));   /* new BitSet(parent.table.size()) (:103) */
    int bad = 0;
    (void)nthreads;
#ifdef _OPENMP
#pragma omp parallel for schedule(static) num_threads(nthreads) if (nthreads > 1)
#endif
    for (int64_t i = 0; i < n->size; i++) {
        if (!bs_get(n->bits, i)) continue;      /* (:106) */
        if (up->akind[i] == 0) continue;        /* Association.None -> no-op (:115-117) */
        for (int64_t e = up->aoff[i]; e < up->aoff[i + 1]; e++) {   /* One (:114) / Many (:111-113) */
            int64_t t = up->atgt[e];
            if (t < 0) { __atomic_store_n(&bad, 1, __ATOMIC_RELAXED); continue; }   /* BitSet.set(negative) -> IndexOutOfBoundsException */
            if (t >= p->size) continue;         /* BitSet auto-grows, the and() below truncates it away */
            if (nthreads > 1) __atomic_fetch_or(&reach[t >> 6], (uint64_t)1 << (t & 63), __ATOMIC_RELAXED);
            else bs_set(reach, t);
        }
    }
    for (int64_t w = 0; w < pw; w++) p->bits[w] &= reach[w];   /* parent.matchingBits.and(...) (:121) */
    free(reach);
    return bad ? ORC_THROW_INDEX_OOB : ORC_SUCCESS;
}

/* ------------------------------------------------------------------ execute (E/DataSystemSerialIndices.java:53-102) */

int orc_execute(orc_system *s, const orc_query *q, int nthreads, uint64_t **out_words, int64_t *out_nwords,
                int32_t **out_indices, int64_t *out_count) {
    if (out_words) *out_words = NULL;
    if (out_indices) *out_indices = NULL;
    if (out_count) *out_count = 0;
    if (out_nwords) *out_nwords = 0;
    s->msg[0] = 0;
    if (nthreads < 1) nthreads = 1;
    int root_table = -1;
    for (int r = 0; r < s->nregs; r++)
        if (strcmp(s->regs[r].name, q->table_name) == 0) root_table = s->regs[r].table;
    if (root_table < 0) {   /* (:54-57) */
        snprintf(s->msg, sizeof s->msg, "The query targets the table '%s' but that table is not registered", q->table_name);
        return ORC_FAILURE;
    }
    xctx x = {0};
    int rc = verify(s, q, root_table, &x);   /* (:61-70) */
    if (rc != ORC_SUCCESS) { xctx_free(&x); return rc; }

    /* phase A (:78-89): nodes.push(root); pop -> filterSelf; leaf ? leaves.push : nodes.addAll(children) */
    int *dq = (int *)xmalloc(sizeof(int) * (size_t)(2 * x.n + 2));
    int *leaves = (int *)xmalloc(sizeof(int) * (size_t)(x.n + 1));
    int dh = x.n, dt = x.n, nl = 0;          /* deque in the middle of the buffer: push=addFirst, addAll=addLast */
    dq[--dh] = 0;
    while (dh < dt) {
        int ni = dq[dh++];
        filter_self(s, &x.nodes[ni], nthreads);
        if (x.nodes[ni].nchildren == 0) leaves[nl++] = ni;
        else for (int k = 0; k < x.nodes[ni].nchildren; k++) dq[dt++] = x.nodes[ni].children[k];
    }
    /* phase B (:92-97): pop a leaf, filterParent, push its parent; each chain runs to the root before the next leaf */
    int *stack = (int *)xmalloc(sizeof(int) * (size_t)(x.n + nl + 2));
    int sp = 0;
    for (int k = 0; k < nl; k++) stack[sp++] = leaves[k];
    while (sp > 0 && rc == ORC_SUCCESS) {
        int ni = stack[--sp];
        rc = filter_parent(s, &x, ni, nthreads);
        if (x.nodes[ni].parent >= 0) stack[sp++] = x.nodes[ni].parent;
    }
    free(dq); free(leaves); free(stack);
    if (rc != ORC_SUCCESS) {
        snprintf(s->msg, sizeof s->msg, "IndexOutOfBoundsException: negative association target");
        xctx_free(&x);
        return rc;
    }

    /* table.subset(executionContext.matchingRows()) (:100): cardinality(), then ascending copy (M/InMemoryTable.java:121-131) */
    xnode *root = &x.nodes[0];
    int64_t nwords = bs_nwords(root->size);
    int64_t card = 0;
    for (int64_t w = 0; w < nwords; w++) card += __builtin_popcountll(root->bits[w]);
    if (out_indices) {
        int32_t *idx = (int32_t *)xmalloc(sizeof(int32_t) * (size_t)card);
        int64_t j = 0;
        for (int64_t w = 0; w < nwords; w++) {
            uint64_t word = root->bits[w];
            while (word) {
                int b = __builtin_ctzll(word);
                idx[j++] = (int32_t)((w << 6) + b);
                word &= word - 1;
            }
        }
        *out_indices = idx;
    }
    if (out_count) *out_count = card;
    if (out_words) {
        *out_words = (uint64_t *)xdup(root->bits, sizeof(uint64_t) * (size_t)nwords);
        if (out_nwords) *out_nwords = nwords;
    }
    s->last_nnodes = x.n < 64 ? x.n : 64;
    for (int i = 0; i < s->last_nnodes; i++) {
        int64_t c = 0;
        for (int64_t w = 0; w < bs_nwords(x.nodes[i].size); w++) c += __builtin_popcountll(x.nodes[i].bits[w]);
        s->last_card[i] = c;
    }
    xctx_free(&x);
    return ORC_SUCCESS;
}

int orc_last_node_cardinalities(const orc_system *s, int64_t *out, int cap) {
    int n = s->last_nnodes < cap ? s->last_nnodes : cap;
    for (int i = 0; i < n; i++) out[i] = s->last_card[i];
    return s->last_nnodes;
}
