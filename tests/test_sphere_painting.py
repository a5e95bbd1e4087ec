"""IONISE_ENTIRE_SPHERE (find_ionised_regions' sphere method, IonisationBox.c:1150-1158;
update_in_sphere / check_region, bubble_helper_progs.c:263-413) against the compiled reference: every
cell within R of a centre that crosses the barrier at radius R is ionised, while only the centres
record z_reion and the ionised-gas temperature.  The product builds the same set as a thresholded
exact distance transform of the centre flags; the ionised mask must be bit-identical."""
import dataclasses

import numpy as np
import pytest

import common

pkg = common.pkg

# (source model, redshift, HII_DIM, NON_CUBIC_FACTOR)
CASES = [("E-INTEGRAL", 8.0, 32, 1.0), ("CONST-ION-EFF", 8.5, 24, 1.0), ("E-INTEGRAL", 8.0, 24, 1.5),
         ("E-INTEGRAL", 10.0, 32, 1.0)]


def _inputs(source, hii, ncf, sphere=True, **astro):
    inp = common.make_inputs(hii=hii, dim=2 * hii, seed=5, source=source, **astro)
    ao = dataclasses.replace(inp.astro_options, IONISE_ENTIRE_SPHERE=sphere)
    so = dataclasses.replace(inp.simulation_options, NON_CUBIC_FACTOR=ncf)
    return dataclasses.replace(inp, astro_options=ao, simulation_options=so)


def _run(be, ref, source, z, hii, ncf):
    inputs = _inputs(source, hii, ncf)
    ics = pkg.compute_initial_conditions(inputs=inputs, backend=ref)
    pf = pkg.perturb_field(redshift=z, initial_conditions=ics, backend=ref)
    got = pkg.compute_ionization_field(perturbed_field=pf, initial_conditions=ics, backend=be)
    want = pkg.compute_ionization_field(perturbed_field=pf, initial_conditions=ics, backend=ref)
    assert common.compare_ionized(got, want)["mask_mismatch"] == 0
    assert np.array_equal(got.neutral_fraction == 0, want.neutral_fraction == 0)
    assert np.array_equal(got.z_reion, want.z_reion)
    # the spheres really add cells: more ionised cells than recorded centres, and more than the centre method
    painted, centres = int((got.neutral_fraction == 0).sum()), int((got.z_reion > 0).sum())
    assert 0 < centres < painted
    plain = _inputs(source, hii, ncf, sphere=False)
    centre_method = pkg.compute_ionization_field(perturbed_field=_retag(pf, plain), initial_conditions=_retag(ics, plain),
                                                 backend=be)
    assert int((centre_method.neutral_fraction == 0).sum()) == centres


def _retag(struct, inputs):
    struct.inputs = inputs
    return struct


@pytest.mark.parametrize("source,z,hii,ncf", CASES)
def test_sphere_painting_emulated_vs_reference(source, z, hii, ncf):
    emu, ref = common.emu_backend(), common.ref_backend()
    if emu is None or ref is None:
        pytest.skip("needs tests/_emu and oracle/_ref")
    _run(emu, ref, source, z, hii, ncf)


@pytest.mark.gpu
@pytest.mark.parametrize("source,z,hii,ncf", [CASES[0], CASES[2]])
def test_sphere_painting_gpu_vs_reference(source, z, hii, ncf):
    ref = common.ref_backend()
    if ref is None:
        pytest.skip("oracle/_ref not present on this box")
    _run(common.gpu_backend(), ref, source, z, hii, ncf)


def test_sphere_painting_refuses_order_dependent_configurations():
    """Where the reference's result depends on the order it visits the cells (spheres larger than a cell
    at the unfiltered radius; spheres combined with recombinations) the library says so (status 3)."""
    emu = common.emu_backend()
    if emu is None:
        pytest.skip("tests/_emu not built")
    inputs = _inputs("E-INTEGRAL", 16, 1.0, R_BUBBLE_MIN=4.0)
    ics = pkg.compute_initial_conditions(inputs=inputs, backend=emu)
    pf = pkg.perturb_field(redshift=8.0, initial_conditions=ics, backend=emu)
    with pytest.raises(pkg.BackendError) as e:
        pkg.compute_ionization_field(perturbed_field=pf, initial_conditions=ics, backend=emu)
    assert e.value.code == 3
    rec = dataclasses.replace(inputs, astro_params=pkg.AstroParams(),
                              astro_options=dataclasses.replace(inputs.astro_options, RECOMB_MODEL="inhomogeneous"))
    with pytest.raises(pkg.BackendError) as e:
        pkg.compute_ionization_field(perturbed_field=_retag(pf, rec), initial_conditions=_retag(ics, rec),
                                     previous_ionized_box=pkg.IonizedBox.initial(rec),
                                     previous_perturbed_field=pkg.PerturbedField.initial(rec), backend=emu)
    assert e.value.code == 3


def test_sphere_painting_ragged_noncubic_emulated():
    """35 x 35 x 42 cells: the sphere's radius is in x-cells while z wraps with its own length
    (update_in_sphere's dimensions / dimensions_ncf), on mixed-radix transforms."""
    emu, ref = common.emu_backend(), common.ref_backend()
    if emu is None or ref is None:
        pytest.skip("needs tests/_emu and oracle/_ref")
    inputs = _inputs("E-INTEGRAL", 35, 1.2)
    ics = pkg.compute_initial_conditions(inputs=inputs, backend=ref)
    pf = pkg.perturb_field(redshift=8.0, initial_conditions=ics, backend=ref)
    got = pkg.compute_ionization_field(perturbed_field=pf, initial_conditions=ics, backend=emu)
    want = pkg.compute_ionization_field(perturbed_field=pf, initial_conditions=ics, backend=ref)
    assert got.neutral_fraction.shape == (35, 35, 42)
    assert np.array_equal(got.neutral_fraction == 0, want.neutral_fraction == 0)
    assert np.array_equal(got.z_reion, want.z_reion)
