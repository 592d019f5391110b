"""``lumol_b200.units`` against the known answers of the reference's own unit tests (lumol-core/src/units.rs:455-494):
every input of the reference's tests and benches reaches the device through these conversion factors."""

import pytest

from lumol_b200 import units


def test_eval_known_answers():
    # units.rs:456-467, exact literals
    assert units.from_(1.0, "A") == 1.0
    assert units.from_(1.0, "nm") == 10.0
    assert units.from_(1.0, "bohr/fs") == 0.52917720859
    assert units.from_(1.0, "(Ry / rad^-3   )") == 0.1312749878912494
    assert units.from_(1.0, "bar/(m * fs^2)") == 6.022141794216763e-19
    assert units.from_(1.0, "kJ/mol/deg^2") == 0.3282806352310398
    assert units.from_(1.0, "(kcal/mol/A)^2") == 1.7505856024515547e-7
    assert abs(units.from_(1.0, "kcal/mol/A^2") - 4.184e-4) < 1e-9


def test_parsing_errors():
    # units.rs:448-453, 470-475
    for text in ("(", ")", "(bar/m", "m/K)", "m^4-8", "foo ^ bar", "m^z4", "HJK"):
        with pytest.raises(ValueError):
            units.from_(1.0, text)


def test_from_str_and_to():
    # units.rs:478-494
    assert units.from_str("10.0 A") == 10.0
    assert units.from_str("10 A") == 10.0
    assert units.from_str("1e1 A") == 10.0
    assert units.from_str("10") == 10.0
    for text in ("10a.0 bar", "h10"):
        with pytest.raises(ValueError):
            units.from_str(text)
    assert units.to(25.0, "m") == 2.5e-9
    assert units.to(25.0, "bar") == 4.1513469550000005e9
    assert units.to(25.0, "kJ/mol") == 249999.99982494753
