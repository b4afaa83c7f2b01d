/*
 * FlatProblem.java -- flattens JAICOV's object graph into the arrays of include/jaicov_b200.h AFTER
 * BundleAdjustment.prepareUnknownParameters() (BundleAdjustment.java:667-782) has assigned rows and columns, and writes the
 * adjusted values back into the UnknownParameters afterwards (what updateUnknownParameters did, :450-462).
 *
 * Source for a maintainer of the reference to add; NOT compiled in this repository's image (no JDK).  The C++ twin of this
 * class is FlatProblem / BundleAdjustment::prepareUnknownParameters in bundle-adjustment_b200/host/jaicov_host.hpp, whose output
 * is checked against the reference's executed bookkeeping (tests/test_host_cpp.py).
 *
 * One accessor has to be added to the reference: DirectlyObservedParameterGroup keeps its dispersion matrix private and inverts
 * it in place in getWeightMatrix() (parameter/DirectlyObservedParameterGroup.java:67-91); the library needs the dispersion itself
 * (packed upper, as handed to the constructor :49-60), so the class gets `double[] getDispersionData()` returning a copy of the
 * packed data taken in the constructor (null for a diagonal model).
 */
package org.applied_geodesy.adjustment.bundle.gpu;

import static java.lang.foreign.ValueLayout.JAVA_BYTE;
import static java.lang.foreign.ValueLayout.JAVA_DOUBLE;
import static java.lang.foreign.ValueLayout.JAVA_INT;
import static java.lang.foreign.ValueLayout.JAVA_LONG;

import java.lang.foreign.Arena;
import java.lang.foreign.MemorySegment;
import java.util.ArrayList;
import java.util.Collection;
import java.util.LinkedHashMap;
import java.util.List;
import java.util.Map;
import java.util.Set;

import org.applied_geodesy.adjustment.bundle.ObjectCoordinate;
import org.applied_geodesy.adjustment.bundle.ScaleBar;
import org.applied_geodesy.adjustment.bundle.camera.Camera;
import org.applied_geodesy.adjustment.bundle.camera.Image;
import org.applied_geodesy.adjustment.bundle.camera.ImageCoordinate;
import org.applied_geodesy.adjustment.bundle.camera.distortion.DistortionModel;
import org.applied_geodesy.adjustment.bundle.parameter.DirectlyObservedParameterGroup;
import org.applied_geodesy.adjustment.bundle.parameter.ObservationParameter;
import org.applied_geodesy.adjustment.bundle.parameter.ParameterType;
import org.applied_geodesy.adjustment.bundle.parameter.PolynomialCoefficient;
import org.applied_geodesy.adjustment.bundle.parameter.UnknownParameter;
import org.applied_geodesy.adjustment.defect.RankDefect;

public final class FlatProblem {
	/** one DirectlyObservedParameterGroup in the layout of jaicov_add_observed_group */
	public static final class Group {
		public int r;
		public MemorySegment kind, index, comp, obs, var, sigmaPacked;   // var or sigmaPacked is MemorySegment.NULL
	}

	public int nCam, nImg, nPt, nBar;
	public long m;
	public MemorySegment ioVal, ioCol, r0, coefPtr, coefType, coefOrder, coefVal, coefCol;
	public MemorySegment camOfImg, eoVal, eoCol, ptPtr;
	public MemorySegment objIdx, xy, var, rho;
	public MemorySegment xyz, ptCol, isDatum;
	public MemorySegment barA, barB, barLen, barVar;
	public final List<Group> groups = new ArrayList<>();

	// the parameters behind the value arrays, in array order, for writeBack()
	private final List<ObjectCoordinate> points = new ArrayList<>();
	private final List<UnknownParameter<?>> ioParams = new ArrayList<>(), coefParams = new ArrayList<>(), eoParams = new ArrayList<>();

	private FlatProblem() {}

	/**
	 * @param cameras                 BundleAdjustment.cameras in their iteration order
	 * @param objectCoordinates       BundleAdjustment.objectCoordinates (the points that take part, first-appearance order)
	 * @param scaleBars               BundleAdjustment.scaleBars
	 * @param observedParameterGroups BundleAdjustment.observedParameterGroups
	 */
	public static FlatProblem of(Collection<Camera> cameras, Set<ObjectCoordinate> objectCoordinates, Collection<ScaleBar> scaleBars,
			Collection<DirectlyObservedParameterGroup> observedParameterGroups, Arena arena) {
		FlatProblem f = new FlatProblem();

		// object points: index = position in objectCoordinates
		Map<ObjectCoordinate, Integer> pointIndex = new LinkedHashMap<>();
		for (ObjectCoordinate oc : objectCoordinates) {
			pointIndex.put(oc, f.points.size());
			f.points.add(oc);
		}
		f.nPt = f.points.size();
		f.xyz = arena.allocate(JAVA_DOUBLE, 3L * Math.max(f.nPt, 1));
		f.ptCol = arena.allocate(JAVA_INT, 3L * Math.max(f.nPt, 1));
		f.isDatum = arena.allocate(JAVA_BYTE, Math.max(f.nPt, 1));
		for (int p = 0; p < f.nPt; p++) {
			ObjectCoordinate oc = f.points.get(p);
			UnknownParameter<?>[] c = {oc.getX(), oc.getY(), oc.getZ()};
			for (int k = 0; k < 3; k++) {
				f.xyz.setAtIndex(JAVA_DOUBLE, 3L * p + k, c[k].getValue());
				f.ptCol.setAtIndex(JAVA_INT, 3L * p + k, c[k].getColumn());
			}
			f.isDatum.set(JAVA_BYTE, p, (byte) (oc.isDatum() ? 1 : 0));
		}

		// cameras: x0, y0, c (iterator order, camera/orientation/InteriorOrientation.java:70-79), then the coefficients of the
		// distortion models in evaluation order (getDistortionModels() is sorted by enum ordinal, camera/Camera.java:50)
		f.nCam = cameras.size();
		List<Image> images = new ArrayList<>();
		List<Integer> camOfImage = new ArrayList<>();
		List<Integer> coefPtr = new ArrayList<>();
		List<Double> r0 = new ArrayList<>();
		coefPtr.add(0);
		Map<UnknownParameter<?>, int[]> slot = new LinkedHashMap<>();   // parameter -> {kind, index, comp} of jaicov_add_observed_group
		int ci = 0;
		for (Camera camera : cameras) {
			int k = 0;
			for (UnknownParameter<?> p : camera.getInteriorOrientation()) {
				slot.put(p, new int[] {1, ci, k++});
				f.ioParams.add(p);
			}
			double camR0 = 0.0;
			for (DistortionModel model : camera.getDistortionModels()) {
				if (model instanceof org.applied_geodesy.adjustment.bundle.camera.distortion.PolynomialDistortionModel pm)
					camR0 = pm.getR0();
				for (UnknownParameter<? extends DistortionModel> p : model) {
					slot.put(p, new int[] {2, f.coefParams.size(), 0});   // position in the GLOBAL coefficient list
					f.coefParams.add(p);
				}
			}
			r0.add(camR0);
			coefPtr.add(f.coefParams.size());
			for (Image image : camera) {
				int e = 0;
				for (UnknownParameter<?> p : image.getExteriorOrientation()) {
					slot.put(p, new int[] {3, images.size(), e++});
					f.eoParams.add(p);
				}
				images.add(image);
				camOfImage.add(ci);
			}
			ci++;
		}
		f.nImg = images.size();
		f.ioVal = doubles(arena, f.ioParams, true);   f.ioCol = ints(arena, f.ioParams);
		f.coefVal = doubles(arena, f.coefParams, true); f.coefCol = ints(arena, f.coefParams);
		f.eoVal = doubles(arena, f.eoParams, true);   f.eoCol = ints(arena, f.eoParams);
		f.r0 = arena.allocate(JAVA_DOUBLE, Math.max(f.nCam, 1));
		f.coefPtr = arena.allocate(JAVA_INT, f.nCam + 1L);
		for (int c = 0; c < f.nCam; c++) f.r0.setAtIndex(JAVA_DOUBLE, c, r0.get(c));
		for (int c = 0; c <= f.nCam; c++) f.coefPtr.setAtIndex(JAVA_INT, c, coefPtr.get(c));
		int nCoef = f.coefParams.size();
		f.coefType = arena.allocate(JAVA_INT, Math.max(nCoef, 1));
		f.coefOrder = arena.allocate(JAVA_INT, Math.max(nCoef, 1));
		for (int c = 0; c < nCoef; c++) {
			UnknownParameter<?> p = f.coefParams.get(c);
			f.coefType.setAtIndex(JAVA_INT, c, p.getParameterType().getId());
			f.coefOrder.setAtIndex(JAVA_INT, c, p instanceof PolynomialCoefficient<?> pc ? pc.getOrder() : 0);
		}

		// images and their observations in row order (rows 2j, 2j+1; :670-676)
		f.camOfImg = arena.allocate(JAVA_INT, Math.max(f.nImg, 1));
		f.ptPtr = arena.allocate(JAVA_LONG, f.nImg + 1L);
		long m = 0;
		for (Image image : images) m += image.getNumberOfImageCoordinates();
		f.m = m;
		f.objIdx = arena.allocate(JAVA_INT, Math.max(m, 1));
		f.xy = arena.allocate(JAVA_DOUBLE, 2 * Math.max(m, 1));
		f.var = arena.allocate(JAVA_DOUBLE, 2 * Math.max(m, 1));
		f.rho = arena.allocate(JAVA_DOUBLE, Math.max(m, 1));
		long j = 0;
		f.ptPtr.setAtIndex(JAVA_LONG, 0, 0L);
		for (int i = 0; i < f.nImg; i++) {
			f.camOfImg.setAtIndex(JAVA_INT, i, camOfImage.get(i));
			for (ImageCoordinate ic : images.get(i)) {
				f.objIdx.setAtIndex(JAVA_INT, j, pointIndex.get(ic.getObjectCoordinate()));
				f.xy.setAtIndex(JAVA_DOUBLE, 2 * j, ic.getX().getValue());
				f.xy.setAtIndex(JAVA_DOUBLE, 2 * j + 1, ic.getY().getValue());
				f.var.setAtIndex(JAVA_DOUBLE, 2 * j, ic.getX().getVariance());
				f.var.setAtIndex(JAVA_DOUBLE, 2 * j + 1, ic.getY().getVariance());
				f.rho.setAtIndex(JAVA_DOUBLE, j, ic.getCorrelationCoefficientXY());
				j++;
			}
			f.ptPtr.setAtIndex(JAVA_LONG, i + 1L, j);
		}

		// scale bars (ScaleBar.java:34-39)
		f.nBar = scaleBars.size();
		f.barA = arena.allocate(JAVA_INT, Math.max(f.nBar, 1));
		f.barB = arena.allocate(JAVA_INT, Math.max(f.nBar, 1));
		f.barLen = arena.allocate(JAVA_DOUBLE, Math.max(f.nBar, 1));
		f.barVar = arena.allocate(JAVA_DOUBLE, Math.max(f.nBar, 1));
		int b = 0;
		for (ScaleBar bar : scaleBars) {
			f.barA.setAtIndex(JAVA_INT, b, pointIndex.get(bar.getObjectCoordinateA()));
			f.barB.setAtIndex(JAVA_INT, b, pointIndex.get(bar.getObjectCoordinateB()));
			f.barLen.setAtIndex(JAVA_DOUBLE, b, bar.getLength().getValue());
			f.barVar.setAtIndex(JAVA_DOUBLE, b, bar.getLength().getVariance());
			b++;
		}

		// directly observed groups (parameter/DirectlyObservedParameterGroup.java:37-105)
		for (DirectlyObservedParameterGroup group : observedParameterGroups) {
			Group g = new Group();
			g.r = group.getNumberOfParameters();
			g.kind = arena.allocate(JAVA_INT, g.r);
			g.index = arena.allocate(JAVA_INT, g.r);
			g.comp = arena.allocate(JAVA_INT, g.r);
			g.obs = arena.allocate(JAVA_DOUBLE, g.r);
			MemorySegment var = arena.allocate(JAVA_DOUBLE, g.r);
			int i = 0;
			for (ObservationParameter<? extends UnknownParameter<?>> op : group) {
				UnknownParameter<?> ref = op.getReference();
				ParameterType t = ref.getParameterType();
				int[] s;
				if (t == ParameterType.OBJECT_COORDINATE_X || t == ParameterType.OBJECT_COORDINATE_Y || t == ParameterType.OBJECT_COORDINATE_Z)
					s = new int[] {0, pointIndex.get((ObjectCoordinate) ref.getReference()), t.getId() - ParameterType.OBJECT_COORDINATE_X.getId()};
				else
					s = slot.get(ref);
				if (s == null)
					throw new IllegalArgumentException("Error, observed parameter does not belong to a camera or image of this adjustment");
				g.kind.setAtIndex(JAVA_INT, i, s[0]);
				g.index.setAtIndex(JAVA_INT, i, s[1]);
				g.comp.setAtIndex(JAVA_INT, i, s[2]);
				g.obs.setAtIndex(JAVA_DOUBLE, i, op.getValue());
				var.setAtIndex(JAVA_DOUBLE, i, op.getVariance());
				i++;
			}
			double[] dispersion = group.getDispersionData();   // the accessor described in the file header; null = diagonal model
			if (dispersion != null) {
				g.sigmaPacked = arena.allocate(JAVA_DOUBLE, dispersion.length);
				MemorySegment.copy(dispersion, 0, g.sigmaPacked, JAVA_DOUBLE, 0, dispersion.length);
				g.var = MemorySegment.NULL;
			} else {
				g.var = var;
				g.sigmaPacked = MemorySegment.NULL;
			}
			f.groups.add(g);
		}
		return f;
	}

	/** free_flags of jaicov_set_datum in the order tx, ty, tz, rx, ry, rz, scale (1 = FREE), defect/RankDefect.java:35-130 */
	public MemorySegment freeFlags(RankDefect rankDefect, Arena arena) {
		boolean[] free = {rankDefect.estimateTranslationX(), rankDefect.estimateTranslationY(), rankDefect.estimateTranslationZ(),
				rankDefect.estimateRotationX(), rankDefect.estimateRotationY(), rankDefect.estimateRotationZ(), rankDefect.estimateScale()};
		MemorySegment s = arena.allocate(JAVA_INT, 7);
		for (int i = 0; i < 7; i++) s.setAtIndex(JAVA_INT, i, free[i] ? 1 : 0);
		return s;
	}

	/** after jaicov_get_values(h, xyz, ioVal, coefVal, eoVal): UnknownParameter.setValue for every parameter (:450-462) */
	public void writeBack() {
		for (int p = 0; p < nPt; p++) {
			ObjectCoordinate oc = points.get(p);
			oc.getX().setValue(xyz.getAtIndex(JAVA_DOUBLE, 3L * p));
			oc.getY().setValue(xyz.getAtIndex(JAVA_DOUBLE, 3L * p + 1));
			oc.getZ().setValue(xyz.getAtIndex(JAVA_DOUBLE, 3L * p + 2));
		}
		for (int i = 0; i < ioParams.size(); i++) ioParams.get(i).setValue(ioVal.getAtIndex(JAVA_DOUBLE, i));
		for (int i = 0; i < coefParams.size(); i++) coefParams.get(i).setValue(coefVal.getAtIndex(JAVA_DOUBLE, i));
		for (int i = 0; i < eoParams.size(); i++) eoParams.get(i).setValue(eoVal.getAtIndex(JAVA_DOUBLE, i));
	}

	private static MemorySegment doubles(Arena arena, List<UnknownParameter<?>> params, boolean values) {
		MemorySegment s = arena.allocate(JAVA_DOUBLE, Math.max(params.size(), 1));
		for (int i = 0; i < params.size(); i++) s.setAtIndex(JAVA_DOUBLE, i, params.get(i).getValue());
		return s;
	}

	private static MemorySegment ints(Arena arena, List<UnknownParameter<?>> params) {
		MemorySegment s = arena.allocate(JAVA_INT, Math.max(params.size(), 1));
		for (int i = 0; i < params.size(); i++) s.setAtIndex(JAVA_INT, i, params.get(i).getColumn());
		return s;
	}
}
