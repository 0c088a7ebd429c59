/* TEST INFRASTRUCTURE ONLY -- plain-C restatement of the two closed-form pieces of the reference
 * step, independent of torch, used by tests/ as a second checker for the CUDA kernels.
 *
 *   oracle_sgd_momentum : optim.SGD(lr, momentum).step()   (cifar_example_ddp.py:86,107;
 *                         torch/optim/sgd.py:_single_tensor_sgd: first step buf = clone(grad),
 *                         later buf = momentum*buf + grad (two roundings), p = p + (-lr)*buf (one FMA))
 *   oracle_ce_loss_grad : nn.CrossEntropyLoss() forward + backward, mean reduction
 *                         (cifar_example_ddp.py:85,105-106; torch/nn/functional.py:3421-3504),
 *                         evaluated in double precision.
 * Build: gcc -O2 -shared -fPIC -ffp-contract=off oracle/step_math.c -o oracle/_build/libstepmath.so -lm
 */
#include <math.h>
#include <stdint.h>

void oracle_sgd_momentum(float *p, const float *g, float *m, int64_t n, float lr, float momentum, int first_step) {
    for (int64_t i = 0; i < n; ++i) {
        float buf;
        if (first_step) buf = g[i];
        else {
            volatile float t = m[i] * momentum;      /* buf.mul_(momentum): rounded product ... */
            buf = t + g[i];                          /* ... .add_(grad): rounded sum            */
        }
        m[i] = buf;
        p[i] = fmaf(-lr, buf, p[i]);                 /* param.add_(buf, alpha=-lr)              */
    }
}

/* logits [B][C], labels [B]; loss_out: mean_b(-log softmax(logits_b)[label_b]); dlogits = (softmax - onehot)/B */
void oracle_ce_loss_grad(const float *logits, const int64_t *labels, int B, int C, double *loss_out, double *dlogits) {
    double total = 0.0;
    for (int b = 0; b < B; ++b) {
        const float *row = logits + (int64_t)b * C;
        double mx = row[0];
        for (int j = 1; j < C; ++j) if (row[j] > mx) mx = row[j];
        double se = 0.0;
        for (int j = 0; j < C; ++j) se += exp((double)row[j] - mx);
        double lse = log(se) + mx;
        total += lse - (double)row[labels[b]];
        if (dlogits)
            for (int j = 0; j < C; ++j)
                dlogits[(int64_t)b * C + j] = (exp((double)row[j] - lse) - (j == labels[b] ? 1.0 : 0.0)) / (double)B;
    }
    *loss_out = total / (double)B;
}
