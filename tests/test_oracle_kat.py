"""Known-answer vectors of SURVEY.md Appendix E (produced by a restatement that reproduces the
goldens) against the oracle: RNG integers, seeding, first event, first batch, whole-run sums."""
import pytest


def test_ranf_first_draws(oracle):
    assert oracle.rng_words("", 0, 12) == [226501086, 568121092, 685283207, 941239788, 187396897, 998530439,
                                            809158933, 653655343, 543220086, 619523769, 202591062, 316026706]


def test_ranf_batch_positions(oracle):
    assert oracle.rng_words("", 1, 4) == [676512371, 251662982, 18403872, 819509718]
    assert oracle.rng_words("", 999, 4) == [706297141, 696929303, 826235218, 443673403]


def test_ranf_jump_reseed(oracle):
    assert oracle.rng_words("multi-threading,faster-threading", 1, 4) == [446722654, 691500292, 294127751, 637030316]


def test_xoshiro256_stream(oracle):
    assert oracle.rng_words("standard-random", 0, 3) == [0x4f2790d70610546a, 0xd2ae33f21d5120ec, 0xa28f6ee203d01e40]


def test_xoshiro128_stream(oracle):
    assert oracle.rng_words("standard-random,f32", 0, 3) == [0xde3fee85, 0xbaa437d0, 0x6da600ec]


def test_event0(oracle, valeurs_text):
    mom, kept, m2 = oracle.events(valeurs_text, "", 1)
    want = [19.745701122938293, 35.225158994824397, -9.8092334700682162, 41.556294352569154,
            -21.921107851913323, -30.446692331689007, 16.617895638383708, 41.032797843237034,
            2.1754067289750281, -4.7784666631353918, -6.8086621683154886, 8.5979078041938006]
    assert mom == pytest.approx(want, rel=1e-13)
    assert kept == [1]
    assert m2 == pytest.approx([0.00060296911301840818, 198.50827287349748, 222.65565794918706, 0.38925059047436861,
                                0.088527010531325001], rel=1e-12)


def test_batch0_and_run_sums(oracle, valeurs_text):
    run = oracle.run(valeurs_text, "", num_events=20000)
    b0 = run.per_batch[0]
    assert b0.selected_events == 7100
    assert b0.sigma * (20000 / 1e7) == pytest.approx(190.10142964705693, rel=1e-12)  # norm_weight ~ 1/N_total
    assert list(b0.spm2) == pytest.approx([3.4542139751480025, 4244257.4596767705, 6520727.2474444089,
                                           2644.1902009258156, 11.365686895820177], rel=1e-12)
    assert list(b0.vars) == pytest.approx([0.0069238092616542993, 2861499847.2049785, 7383663346.9659595,
                                           1163.9390466227383, 170.58315130676095], rel=1e-12)
    # sequential scheduler: first batch, full batches, then a (possibly empty) remainder (sequential.rs:24-36)
    assert len(run.per_batch) == 3 and run.per_batch[2].selected_events == 0


def test_constants(oracle, valeurs_text):
    c = oracle.constants(valeurs_text)
    assert c[0] == pytest.approx(-0.02776916595454328, rel=1e-15)
    assert c[1] == pytest.approx(-5.3688217230532914e-09, rel=1e-15)
    assert c[3] == pytest.approx(10258.305161475482, rel=1e-15)
    assert c[4] == pytest.approx(1.0475536451880148e-07, rel=1e-15)
    assert c[5:] == pytest.approx([3.3991340182922487, 1.656968561715729e-05, 1.656968561715729e-05, 0.0,
                                   -0.0010838634696015194], rel=1e-14)
