"""The API cases of tests/api_cases.py on the CUDA kernels (vectors from the real reference)."""
import pytest

from api_cases import CASES, load_api_golden

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def api_golden():
    return load_api_golden()


@pytest.mark.parametrize("case", CASES, ids=lambda f: f.__name__)
def test_api_case_on_device(case, api_golden):
    import symmer_b200
    from symmer_b200 import ops
    before = ops.launch_count()
    case(symmer_b200, api_golden)
    assert ops.launch_count() > before or case.__name__ == "case_misc_methods"
