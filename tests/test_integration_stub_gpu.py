"""The ctypes stub printed in INTEGRATION.md section 2 is executed as written (only the library path
is made absolute) and must give the columns of the package's own binding."""
import os
import re

import numpy as np
import pytest
import torch

from nanocompore_b200 import _lib as L
from nanocompore_b200.engine import make_labels
from nanocompore_b200.synthetic import make_transcript
from oracle import prep as oprep

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_integration_stub_runs_as_documented(engine_factory):
    text = open(os.path.join(ROOT, 'INTEGRATION.md')).read()
    code = re.search(r"```python\n# nanocompore/_ncb.py\n(.*?)```", text, flags=re.S).group(1)
    code = code.replace("C.CDLL('libncb.so')", f"C.CDLL({L.LIB_PATH!r})")
    ns = {}
    exec(compile(code, 'INTEGRATION.md', 'exec'), ns)

    tx = make_transcript(np.random.default_rng(77), 120, 60)
    data, samples, conditions, _ = oprep.prepare_data(tx.data, tx.samples, tx.conditions)
    dev = torch.from_numpy(data).cuda()
    handle = ns['make_handle'](0)
    status, pv, st, cnt = ns['compare'](handle, dev, torch.from_numpy(samples.astype(np.int64)).cuda(),
                                        torch.from_numpy(conditions.astype(np.int64)).cuda(), 4)
    ns['lib'].ncb_destroy(handle)

    eng = engine_factory(tests=('GMM', 'KS'))
    want = eng.compare_dense(dev, torch.from_numpy(make_labels(samples, conditions)).cuda()).numpy()
    np.testing.assert_array_equal(status.cpu().numpy(), want['status'])
    np.testing.assert_array_equal(pv.cpu().numpy(), want['pvalues'])
    np.testing.assert_array_equal(st.cpu().numpy(), want['stats'])
    np.testing.assert_array_equal(cnt.cpu().numpy(), want['counts'])
    assert np.isfinite(want['pvalues'][L.PV['KS_INTENSITY']]).sum() > 50
