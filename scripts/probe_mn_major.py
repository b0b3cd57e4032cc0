import sys; sys.path.insert(0, '.')
from tests.test_umma_gpu import _run_mn
for order in (0, 1):
    for n, k in ((16, 16), (32, 64), (64, 128)):
        try:
            print('order', order, n, k, _run_mn(n, k, order))
        except Exception as e:
            print('order', order, n, k, 'ERR', e)
for shift in (1, 3, 8):
    print('shift', shift, _run_mn(32, 160, 0, shift=shift, use=128), _run_mn(32, 160, 1, shift=shift, use=128))
