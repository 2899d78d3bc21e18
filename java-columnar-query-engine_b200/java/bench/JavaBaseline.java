package dgroomes.bench;

import dgroomes.data_system.Association;
import dgroomes.data_system.Criteria;
import dgroomes.data_system.Query;
import dgroomes.data_system.QueryResult;
import dgroomes.data_system_serial_indices_arrays.DataSystemSerialIndices;
import dgroomes.geography.City;
import dgroomes.geography.GeographyGraph;
import dgroomes.geography.State;
import dgroomes.geography.Zip;
import dgroomes.geography_loader.GeographiesLoader;
import dgroomes.geography_loader.StateData;
import dgroomes.in_memory.InMemoryColumn;
import dgroomes.in_memory.InMemoryTable;

import java.io.File;
import java.util.ArrayList;
import java.util.Arrays;
import java.util.HashMap;
import java.util.List;
import java.util.Map;

/**
 * Times the REFERENCE engine (DataSystemSerialIndices.execute) on the workload bench.py measures: the Plymouth-adjacency
 * query (app/.../Runner.java:230-236) over U "parallel universes" (SURVEY.md 8d: ZIP and city rows replicated U times with
 * foreign keys rebased per universe, one shared states table).
 * <p>
 * SOURCE ONLY: neither the build image nor the GPU box of this project has a JVM, so this file was never compiled or
 * run there; bench.py times a C port of the same algorithm instead and says so in its JSON line.  On a machine with
 * JDK 21+: add this file to a module that depends on :app's dependencies and run
 * {@code java dgroomes.bench.JavaBaseline ../zips.jsonl 1000 10}; it prints ZIP rows per second like bench.py does.
 */
public final class JavaBaseline {

    public static void main(String[] args) {
        File zipsFile = new File(args.length > 0 ? args[0] : "../zips.jsonl");
        int universes = args.length > 1 ? Integer.parseInt(args[1]) : 1000;
        int runs = args.length > 2 ? Integer.parseInt(args[2]) : 10;

        GeographyGraph geo = GeographiesLoader.loadFromFile(zipsFile);
        List<State> states = new ArrayList<>(geo.states());
        List<City> cities = new ArrayList<>(geo.cities());
        List<Zip> zips = new ArrayList<>(geo.zips());
        Map<State, Integer> stateIndex = new HashMap<>();
        Map<City, Integer> cityIndex = new HashMap<>();
        for (int i = 0; i < states.size(); i++) stateIndex.put(states.get(i), i);
        for (int i = 0; i < cities.size(); i++) cityIndex.put(cities.get(i), i);

        // states: shared by every universe
        String[] stateCodes = new String[states.size()], stateNames = new String[states.size()];
        for (int i = 0; i < states.size(); i++) { stateCodes[i] = states.get(i).code(); stateNames[i] = states.get(i).name(); }
        InMemoryTable statesTable = InMemoryTable.ofColumns(new InMemoryColumn.StringColumn(stateCodes), new InMemoryColumn.StringColumn(stateNames));

        // cities and zips: universe u is an exact copy with the ZIP -> city key rebased by u * |cities|
        int nc = cities.size(), nz = zips.size();
        String[] cityNames = new String[nc * universes];
        Association[] cityState = new Association[nc * universes];
        int[] codes = new int[nz * universes], pops = new int[nz * universes];
        Association[] zipCity = new Association[nz * universes];
        for (int u = 0; u < universes; u++) {
            for (int c = 0; c < nc; c++) {
                cityNames[u * nc + c] = cities.get(c).name();
                cityState[u * nc + c] = new Association.One(stateIndex.get(cities.get(c).state(geo)));
            }
            for (int z = 0; z < nz; z++) {
                codes[u * nz + z] = zips.get(z).zipCode();
                pops[u * nz + z] = zips.get(z).population();
                zipCity[u * nz + z] = new Association.One(u * nc + cityIndex.get(zips.get(z).city(geo)));
            }
        }
        InMemoryTable citiesTable = InMemoryTable.ofColumns(new InMemoryColumn.StringColumn(cityNames));
        InMemoryTable zipsTable = InMemoryTable.ofColumns(new InMemoryColumn.IntegerColumn(codes), new InMemoryColumn.IntegerColumn(pops));

        DataSystemSerialIndices dataSystem = new DataSystemSerialIndices();
        dataSystem.register("states", statesTable);
        dataSystem.register("cities", citiesTable);
        dataSystem.register("zips", zipsTable);
        citiesTable.associateTo(statesTable, cityState);      // cities: [0 name, 1 ->state, 2 <-zips]; states: [.., 2 <-cities]
        zipsTable.associateTo(citiesTable, zipCity);          // zips:   [0 code, 1 population, 2 ->city]
        Map<String, State> byCode = new HashMap<>();
        for (State s : states) byCode.put(s.code(), s);
        Association[] adjacency = new Association[states.size()];
        Arrays.fill(adjacency, Association.NONE);
        for (StateData.StateAdjacency a : StateData.STATE_ADJACENCIES) {
            int from = stateIndex.get(byCode.get(a.state())), to = stateIndex.get(byCode.get(a.adjacentState()));
            adjacency[from] = adjacency[from].add(to);
        }
        statesTable.associateTo(statesTable, adjacency);      // states: [.., 3 ->adjacent, 4 <-adjacent]

        long expected = -1;
        double best = Double.MAX_VALUE, sum = 0;
        for (int run = 0; run < runs + 3; run++) {            // 3 warm-ups for the JIT
            Query query = new Query("zips");                  // Runner.java:230-236
            query.rootNode.addCriteria(new Criteria.IntCriteria(1, i -> i >= 10_000 && i < 10_100));
            query.rootNode.createChild(2).createChild(1).createChild(3).createChild(2)
                    .addCriteria(new Criteria.StringCriteria(0, "PLYMOUTH"::equals));
            long t0 = System.nanoTime();
            QueryResult result = dataSystem.execute(query);
            double seconds = (System.nanoTime() - t0) * 1e-9;
            if (!(result instanceof QueryResult.Success(var table))) throw new IllegalStateException(result.toString());
            if (expected < 0) expected = table.size();
            if (table.size() != expected) throw new IllegalStateException("result size changed between runs");
            if (run >= 3) { best = Math.min(best, seconds); sum += seconds; }
        }
        long rows = (long) nz * universes;
        System.out.printf("{\"impl\": \"reference-java\", \"universes\": %d, \"zip_rows\": %d, \"matches\": %d, \"threads\": 1, " +
                        "\"mean_ms\": %.3f, \"best_ms\": %.3f, \"rows_per_s_mean\": %.0f}%n",
                universes, rows, expected, sum / runs * 1e3, best * 1e3, rows / (sum / runs));
    }
}
