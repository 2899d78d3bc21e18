// Sibling of data-system-serial-indices-arrays (add `include("data-system-b200")` to settings.gradle.kts:3-10).
// Needs JDK 22+ (java.lang.foreign is final since JEP 454). NOT compiled in the build image: there is no JVM there.
plugins {
    id("dgroomes.conventions")
    `java-library`
}

dependencies {
    api(project(":data-system"))
    // the shim reads column arrays through the InMemoryColumn record accessors (Column exposes no raw accessor)
    implementation(project(":data-model-in-memory"))
    // the TCK also runs against the reference's own engine (QueryTckTest.ReferenceSerialIndices)
    testImplementation(project(":data-system-serial-indices-arrays"))
    testImplementation(libs.junit.jupiter.api)
    testImplementation(libs.assertj)
    testRuntimeOnly(libs.junit.jupiter.engine)
}

tasks.withType<Test> {
    // where lib/libcolq.so was built
    systemProperty("colq.library", System.getProperty("colq.library") ?: "libcolq.so")
    systemProperty("colq.gpus", System.getProperty("colq.gpus") ?: "2")   // GPUs QueryTckTest.AllGpusOneJvm drives from this JVM
    jvmArgs("--enable-native-access=dgroomes.data_system_b200")
}
