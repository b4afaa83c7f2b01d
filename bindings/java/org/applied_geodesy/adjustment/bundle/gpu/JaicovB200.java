/*
 * JaicovB200.java -- Panama FFM binding of libjaicov_b200.so (include/jaicov_b200.h) for JAICOV.
 *
 * Source for a maintainer of the reference to add (package org.applied_geodesy.adjustment.bundle.gpu; the reference already needs
 * Java 25, parameter/ObservationParameter.java:36-41, so java.lang.foreign is available).  NOT compiled in this repository's
 * image: it holds no JDK.  The same call sequence is exercised by the C++ twin (bundle-adjustment_b200/host/jaicov_host.hpp:
 * BundleAdjustment::estimateModel) and by the Python ctypes binding (bundle-adjustment_b200/_lib.py); the struct offsets used here
 * are pinned by tests/test_lib_symbols.py::test_struct_offsets_used_by_the_java_binding.
 */
package org.applied_geodesy.adjustment.bundle.gpu;

import static java.lang.foreign.ValueLayout.ADDRESS;
import static java.lang.foreign.ValueLayout.JAVA_DOUBLE;
import static java.lang.foreign.ValueLayout.JAVA_INT;
import static java.lang.foreign.ValueLayout.JAVA_LONG;

import java.lang.foreign.Arena;
import java.lang.foreign.FunctionDescriptor;
import java.lang.foreign.Linker;
import java.lang.foreign.MemorySegment;
import java.lang.foreign.SymbolLookup;
import java.lang.invoke.MethodHandle;

public final class JaicovB200 {
	private JaicovB200() {}

	/** EstimationStateType ids the library returns (adjustment/EstimationStateType.java:25-41) plus its own argument error. */
	public static final int OK = 0, ILLEGAL_ARGUMENT = -100;

	/** sizeof(jaicov_options), sizeof(jaicov_stats) */
	public static final long OPTIONS_BYTES = 48, STATS_BYTES = 112;
	/** byte offset of jaicov_options.n_devices: 1 = one GPU; k > 1 = this ONE handle (and the one JVM thread calling it) drives k GPUs */
	public static final long OPTIONS_N_DEVICES = 28;

	private static final Linker LINKER = Linker.nativeLinker();
	private static final SymbolLookup LIB = SymbolLookup.libraryLookup("libjaicov_b200.so", Arena.global());

	private static MethodHandle h(String name, FunctionDescriptor d) {
		return LINKER.downcallHandle(LIB.find(name).orElseThrow(), d);
	}

	public static final MethodHandle DEFAULT_OPTIONS = h("jaicov_default_options", FunctionDescriptor.of(JAVA_INT, ADDRESS));
	public static final MethodHandle CREATE    = h("jaicov_create", FunctionDescriptor.of(JAVA_INT, ADDRESS, ADDRESS));
	public static final MethodHandle DESTROY   = h("jaicov_destroy", FunctionDescriptor.ofVoid(ADDRESS));
	public static final MethodHandle LAST_ERROR = h("jaicov_last_error", FunctionDescriptor.of(ADDRESS, ADDRESS));
	public static final MethodHandle SET_CAMS  = h("jaicov_set_cameras", FunctionDescriptor.of(JAVA_INT, ADDRESS, JAVA_INT, ADDRESS, ADDRESS, ADDRESS, ADDRESS, ADDRESS, ADDRESS, ADDRESS, ADDRESS));
	public static final MethodHandle SET_IMGS  = h("jaicov_set_images", FunctionDescriptor.of(JAVA_INT, ADDRESS, JAVA_INT, ADDRESS, ADDRESS, ADDRESS, ADDRESS));
	public static final MethodHandle SET_OBS   = h("jaicov_set_image_points", FunctionDescriptor.of(JAVA_INT, ADDRESS, JAVA_LONG, ADDRESS, ADDRESS, ADDRESS, ADDRESS));
	public static final MethodHandle SET_PTS   = h("jaicov_set_object_points", FunctionDescriptor.of(JAVA_INT, ADDRESS, JAVA_INT, ADDRESS, ADDRESS, ADDRESS));
	public static final MethodHandle SET_BARS  = h("jaicov_set_scale_bars", FunctionDescriptor.of(JAVA_INT, ADDRESS, JAVA_INT, ADDRESS, ADDRESS, ADDRESS, ADDRESS));
	public static final MethodHandle ADD_GROUP = h("jaicov_add_observed_group", FunctionDescriptor.of(JAVA_INT, ADDRESS, JAVA_INT, ADDRESS, ADDRESS, ADDRESS, ADDRESS, ADDRESS, ADDRESS));
	public static final MethodHandle SET_DATUM = h("jaicov_set_datum", FunctionDescriptor.of(JAVA_INT, ADDRESS, ADDRESS, JAVA_INT, JAVA_INT));
	public static final MethodHandle SET_REDUCED_ROWS = h("jaicov_set_reduced_rows", FunctionDescriptor.of(JAVA_INT, ADDRESS, JAVA_INT));
	public static final MethodHandle ESTIMATE  = h("jaicov_estimate", FunctionDescriptor.of(JAVA_INT, ADDRESS, ADDRESS, ADDRESS, ADDRESS));
	public static final MethodHandle GET_STATS = h("jaicov_get_stats", FunctionDescriptor.of(JAVA_INT, ADDRESS, ADDRESS));
	public static final MethodHandle GET_VALUES = h("jaicov_get_values", FunctionDescriptor.of(JAVA_INT, ADDRESS, ADDRESS, ADDRESS, ADDRESS, ADDRESS));
	public static final MethodHandle GET_QXX_PACKED = h("jaicov_get_qxx_packed", FunctionDescriptor.of(JAVA_INT, ADDRESS, ADDRESS));
	public static final MethodHandle GET_QXX_BLOCK = h("jaicov_get_qxx_block", FunctionDescriptor.of(JAVA_INT, ADDRESS, JAVA_INT, JAVA_INT, JAVA_INT, JAVA_INT, ADDRESS, JAVA_LONG));
	/** optional self-check after estimateModel(): y = K x matrix-free from the observations (K Qxx e_c = e_c on a few columns) */
	public static final MethodHandle NORMAL_PRODUCT = h("jaicov_normal_product", FunctionDescriptor.of(JAVA_INT, ADDRESS, JAVA_INT, ADDRESS, ADDRESS, ADDRESS, ADDRESS));
	/** extension: fully populated dispersion of the image coordinates of one image (not in the reference) */
	public static final MethodHandle SET_IMAGE_DISPERSION = h("jaicov_set_image_dispersion", FunctionDescriptor.of(JAVA_INT, ADDRESS, JAVA_INT, JAVA_LONG, ADDRESS));
	public static final MethodHandle GET_QXX_SUBMATRIX = h("jaicov_get_qxx_submatrix", FunctionDescriptor.of(JAVA_INT, ADDRESS, JAVA_INT, ADDRESS, JAVA_DOUBLE, ADDRESS));

	/** progress callback: void (*)(void *user, int32_t state, double old_value, double new_value) -> PropertyChangeSupport.firePropertyChange */
	public static final FunctionDescriptor PROGRESS_CALLBACK = FunctionDescriptor.ofVoid(ADDRESS, JAVA_INT, JAVA_DOUBLE, JAVA_DOUBLE);

	/** Throws what the reference would throw for the library's non-state return codes. */
	public static void check(int rc, MemorySegment handle) throws Throwable {
		if (rc == OK)
			return;
		String msg = "jaicov_b200 error " + rc;
		if (handle != null && !handle.equals(MemorySegment.NULL)) {
			MemorySegment text = ((MemorySegment) LAST_ERROR.invokeExact(handle)).reinterpret(1024);
			msg += ": " + text.getString(0);
		}
		if (rc == ILLEGAL_ARGUMENT)
			throw new IllegalArgumentException(msg);
		throw new IllegalStateException(msg);
	}
}
