package dgroomes.data_system_b200;

import java.nio.charset.StandardCharsets;
import java.util.function.IntPredicate;
import java.util.function.Predicate;

/**
 * Structured predicates. They implement the same functional interfaces the reference's criteria records hold
 * (data-system/.../Criteria.java:17-19), so they fit an unchanged {@code Criteria.IntCriteria} /
 * {@code Criteria.StringCriteria} and still work on the serial engine; {@link DataSystemColq} recognises them with
 * {@code instanceof} and ships their constants to the GPU. An opaque lambda cannot run on a GPU and yields
 * {@code QueryResult.Failure} (there is no CPU fallback).
 * <p>
 * The five lambda sites of app/.../Runner.java become:
 * <pre>
 *   :231  i -> i >= 10_000 && i < 10_100      ->  Predicates.intHalfOpen(10_000, 10_100)
 *   :236  "PLYMOUTH"::equals                  ->  Predicates.strEquals("PLYMOUTH")
 *   :255  s -> s.contains("North")            ->  Predicates.strContains("North")   (likewise :257, :259)
 * </pre>
 */
public final class Predicates {

    private Predicates() {}

    /** Closed interval {@code lo <= v <= hi}. */
    public record IntRange(int lo, int hi) implements IntPredicate {
        @Override
        public boolean test(int v) {
            return lo <= v && v <= hi;
        }
    }

    /** Operator codes: the {@code colq_str_op} enum of include/colq.h. */
    public enum StrOp {
        EQ, CONTAINS, CMP_GT, CMP_LT, CMP_GE, CMP_LE, NE, STARTS_WITH, ENDS_WITH
    }

    public record StringOp(StrOp op, String value) implements Predicate<String> {
        @Override
        public boolean test(String s) {
            return switch (op) {
                case EQ -> s.equals(value);
                case NE -> !s.equals(value);
                case CONTAINS -> s.contains(value);
                case CMP_GT -> s.compareTo(value) > 0;
                case CMP_LT -> s.compareTo(value) < 0;
                case CMP_GE -> s.compareTo(value) >= 0;
                case CMP_LE -> s.compareTo(value) <= 0;
                case STARTS_WITH -> s.startsWith(value);
                case ENDS_WITH -> s.endsWith(value);
            };
        }

        byte[] needle() {
            return value.getBytes(StandardCharsets.UTF_8);
        }
    }

    public static IntRange intRange(int lo, int hi) { return new IntRange(lo, hi); }
    public static IntRange intHalfOpen(int lo, int hiExclusive) { return new IntRange(lo, hiExclusive - 1); }
    public static IntRange intGreaterThan(int x) { return new IntRange(x == Integer.MAX_VALUE ? 1 : x + 1, x == Integer.MAX_VALUE ? 0 : Integer.MAX_VALUE); }
    public static IntRange intLessThan(int x) { return new IntRange(x == Integer.MIN_VALUE ? 1 : Integer.MIN_VALUE, x == Integer.MIN_VALUE ? 0 : x - 1); }
    public static IntRange intBetweenExclusive(int lo, int hi) { return new IntRange(lo + 1, hi - 1); }
    public static StringOp strEquals(String x) { return new StringOp(StrOp.EQ, x); }
    public static StringOp strContains(String x) { return new StringOp(StrOp.CONTAINS, x); }
    public static StringOp strCompareGt(String x) { return new StringOp(StrOp.CMP_GT, x); }
    public static StringOp strCompareLt(String x) { return new StringOp(StrOp.CMP_LT, x); }
}
