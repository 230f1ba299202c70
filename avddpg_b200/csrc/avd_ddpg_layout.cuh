// avd_ddpg_layout.cuh -- offsets of the named tensors inside the flat per-agent parameter vectors
// (see include/avddpg_b200.h, "DDPG agents").  Trainable tensors come first, in the order of Keras
// `trainable_variables` for agent/model.py's functional models; BN moving statistics form the tail.
#pragma once
#include <stdint.h>

#include "../../include/avddpg_b200.h"

namespace avd {

struct ActorOff {
    int64_t W1, b1, g1, be1, W2, b2, g2, be2, W3, b3, n_train, mu1, var1, mu2, var2, total;
};
struct CriticOff {
    int64_t Ws, bs, Wa, ba, gs, bes, ga, bea, W2, b2, g2, be2, W3, b3, n_train, mus, vars, mua, vara, mu2, var2, total;
};

__host__ __device__ inline ActorOff actor_off(const avd_net_dims& d) {
    ActorOff o;
    int64_t p = 0;
    o.W1 = p; p += (int64_t)d.ns * d.l1;
    o.b1 = p; p += d.l1;
    o.g1 = p; p += d.l1;
    o.be1 = p; p += d.l1;
    o.W2 = p; p += (int64_t)d.l1 * d.l2;
    o.b2 = p; p += d.l2;
    o.g2 = p; p += d.l2;
    o.be2 = p; p += d.l2;
    o.W3 = p; p += d.l2;
    o.b3 = p; p += 1;
    o.n_train = p;
    o.mu1 = p; p += d.l1;
    o.var1 = p; p += d.l1;
    o.mu2 = p; p += d.l2;
    o.var2 = p; p += d.l2;
    o.total = p;
    return o;
}

__host__ __device__ inline CriticOff critic_off(const avd_net_dims& d) {
    CriticOff o;
    int64_t p = 0;
    o.Ws = p; p += (int64_t)d.ns * d.l1;
    o.bs = p; p += d.l1;
    o.Wa = p; p += d.la;
    o.ba = p; p += d.la;
    o.gs = p; p += d.l1;
    o.bes = p; p += d.l1;
    o.ga = p; p += d.la;
    o.bea = p; p += d.la;
    o.W2 = p; p += (int64_t)(d.l1 + d.la) * d.l2;
    o.b2 = p; p += d.l2;
    o.g2 = p; p += d.l2;
    o.be2 = p; p += d.l2;
    o.W3 = p; p += d.l2;
    o.b3 = p; p += 1;
    o.n_train = p;
    o.mus = p; p += d.l1;
    o.vars = p; p += d.l1;
    o.mua = p; p += d.la;
    o.vara = p; p += d.la;
    o.mu2 = p; p += d.l2;
    o.var2 = p; p += d.l2;
    o.total = p;
    return o;
}

constexpr float kBnEps = 1e-3f;  // tf.keras.layers.BatchNormalization default epsilon

}  // namespace avd
