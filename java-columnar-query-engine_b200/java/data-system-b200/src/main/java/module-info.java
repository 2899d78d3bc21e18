module dgroomes.data_system_b200 {
    requires transitive dgroomes.data_system;
    requires dgroomes.in_memory;
    exports dgroomes.data_system_b200;
}
