package dgroomes.data_system_b200;

import java.lang.foreign.Arena;
import java.lang.foreign.FunctionDescriptor;
import java.lang.foreign.Linker;
import java.lang.foreign.MemoryLayout;
import java.lang.foreign.SymbolLookup;
import java.lang.invoke.MethodHandle;

import static java.lang.foreign.ValueLayout.ADDRESS;
import static java.lang.foreign.ValueLayout.JAVA_INT;
import static java.lang.foreign.ValueLayout.JAVA_LONG;

/**
 * java.lang.foreign downcall handles for libcolq.so: one handle per symbol of include/colq.h, no glue logic.
 * (Only the symbols the shim uses are bound; the rest of the header binds the same way.)
 */
final class ColqLibrary {

    static final int OK = 0, FAILURE = 1, THROW_INDEX_OOB = 2, THROW_NULL = 3, THROW_ILLEGAL_STATE = 4,
            THROW_ILLEGAL_ARG = 5, ERR_DEVICE = 6, ERR_CAPACITY = 7;
    static final int REPLICATED = 0, SHARDED = 1;

    private static final Linker LINKER = Linker.nativeLinker();
    private static final SymbolLookup LOOKUP =
            SymbolLookup.libraryLookup(System.getProperty("colq.library", "libcolq.so"), Arena.global());

    private static MethodHandle h(String name, MemoryLayout res, MemoryLayout... args) {
        return LINKER.downcallHandle(LOOKUP.find(name).orElseThrow(() -> new UnsatisfiedLinkError(name)),
                FunctionDescriptor.of(res, args));
    }

    static final MethodHandle colq_create = h("colq_create", JAVA_INT, JAVA_INT, ADDRESS);
    static final MethodHandle colq_destroy = h("colq_destroy", JAVA_INT, ADDRESS);
    static final MethodHandle colq_last_error = h("colq_last_error", ADDRESS, ADDRESS);
    static final MethodHandle colq_table_create = h("colq_table_create", JAVA_INT, ADDRESS, JAVA_LONG, JAVA_INT, JAVA_LONG, ADDRESS);
    static final MethodHandle colq_register = h("colq_register", JAVA_INT, ADDRESS, ADDRESS, JAVA_INT);
    static final MethodHandle colq_col_i32 = h("colq_col_i32", JAVA_INT, ADDRESS, JAVA_INT, JAVA_INT, ADDRESS, JAVA_LONG);
    static final MethodHandle colq_col_str = h("colq_col_str", JAVA_INT, ADDRESS, JAVA_INT, JAVA_INT, ADDRESS, ADDRESS, JAVA_LONG, JAVA_LONG);
    static final MethodHandle colq_col_bool = h("colq_col_bool", JAVA_INT, ADDRESS, JAVA_INT, JAVA_INT, ADDRESS, JAVA_LONG);
    static final MethodHandle colq_associate_fk = h("colq_associate_fk", JAVA_INT, ADDRESS, JAVA_INT, JAVA_INT, JAVA_INT, JAVA_INT, ADDRESS, JAVA_LONG);
    static final MethodHandle colq_associate_csr = h("colq_associate_csr", JAVA_INT, ADDRESS, JAVA_INT, JAVA_INT, JAVA_INT, JAVA_INT, ADDRESS, ADDRESS, JAVA_LONG, JAVA_LONG);
    static final MethodHandle colq_query_create = h("colq_query_create", JAVA_INT, ADDRESS, ADDRESS, ADDRESS);
    static final MethodHandle colq_query_destroy = h("colq_query_destroy", JAVA_INT, ADDRESS);
    static final MethodHandle colq_query_child = h("colq_query_child", JAVA_INT, ADDRESS, JAVA_INT, JAVA_INT, ADDRESS);
    static final MethodHandle colq_query_criteria_i32_range = h("colq_query_criteria_i32_range", JAVA_INT, ADDRESS, JAVA_INT, JAVA_INT, JAVA_INT, JAVA_INT);
    static final MethodHandle colq_query_criteria_str = h("colq_query_criteria_str", JAVA_INT, ADDRESS, JAVA_INT, JAVA_INT, JAVA_INT, ADDRESS, JAVA_INT);
    // host-resident columns: pinned, device-mapped off-heap buffers the kernels read in place (include/colq.h)
    static final MethodHandle colq_host_alloc = h("colq_host_alloc", JAVA_INT, ADDRESS, JAVA_LONG, ADDRESS);
    static final MethodHandle colq_host_free = h("colq_host_free", JAVA_INT, ADDRESS, ADDRESS);
    static final MethodHandle colq_col_i32_host = h("colq_col_i32_host", JAVA_INT, ADDRESS, JAVA_INT, JAVA_INT, ADDRESS, JAVA_LONG, JAVA_LONG);
    static final MethodHandle colq_col_str_host = h("colq_col_str_host", JAVA_INT, ADDRESS, JAVA_INT, JAVA_INT, ADDRESS, JAVA_LONG, ADDRESS, JAVA_LONG, JAVA_LONG, JAVA_LONG);
    static final MethodHandle colq_associate_fk_host = h("colq_associate_fk_host", JAVA_INT, ADDRESS, JAVA_INT, JAVA_INT, JAVA_INT, JAVA_INT, ADDRESS, JAVA_LONG, JAVA_LONG);
    // dictionary-encoded string columns and opaque Predicate<String> criteria evaluated per distinct value
    static final MethodHandle colq_col_str_dict = h("colq_col_str_dict", JAVA_INT, ADDRESS, JAVA_INT, JAVA_INT, ADDRESS, JAVA_LONG, ADDRESS, ADDRESS, JAVA_LONG, JAVA_LONG);
    static final MethodHandle colq_col_str_dict_host = h("colq_col_str_dict_host", JAVA_INT, ADDRESS, JAVA_INT, JAVA_INT, ADDRESS, JAVA_LONG, JAVA_LONG, ADDRESS, ADDRESS, JAVA_LONG, JAVA_LONG);
    static final MethodHandle colq_col_i32_dict = h("colq_col_i32_dict", JAVA_INT, ADDRESS, JAVA_INT, JAVA_INT, ADDRESS, JAVA_LONG, ADDRESS, JAVA_LONG);
    static final MethodHandle colq_col_i32_dict_host = h("colq_col_i32_dict_host", JAVA_INT, ADDRESS, JAVA_INT, JAVA_INT, ADDRESS, JAVA_LONG, JAVA_LONG, ADDRESS, JAVA_LONG);
    static final MethodHandle colq_query_criteria_i32_accept = h("colq_query_criteria_i32_accept", JAVA_INT, ADDRESS, JAVA_INT, JAVA_INT, ADDRESS, JAVA_LONG);
    static final MethodHandle colq_query_criteria_str_accept = h("colq_query_criteria_str_accept", JAVA_INT, ADDRESS, JAVA_INT, JAVA_INT, ADDRESS, JAVA_LONG);
    static final MethodHandle colq_execute = h("colq_execute", JAVA_INT, ADDRESS, ADDRESS, ADDRESS, JAVA_LONG, ADDRESS, JAVA_LONG, ADDRESS, ADDRESS);

    private ColqLibrary() {}
}
