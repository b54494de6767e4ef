// Writes tests/golden/varstore_libtorch.ot with libtorch's own serializer — the code path
// tch 0.3.0's VarStore::save takes (at_save_multi -> torch::serialize::OutputArchive::write per
// variable, then save_to).  Built against the libtorch that ships inside the torch wheel:
//   sh tests/golden/make_varstore_fixture.sh
// The fixture pins csrc/varstore.cu (native reader) to a file libtorch really wrote.
#include <torch/torch.h>

int main(int argc, char **argv) {
  torch::manual_seed(7);
  torch::serialize::OutputArchive ar;
  // VarStore-style names, small shapes
  ar.write("conv1.weight", torch::randn({4, 1, 7, 7}));
  ar.write("bn1.weight", torch::rand({4}));
  ar.write("bn1.bias", torch::zeros({4}));
  ar.write("bn1.running_mean", torch::randn({4}) * 0.1);
  ar.write("bn1.running_var", torch::rand({4}) + 0.5);
  ar.write("layer2.0.downsample.0.weight", torch::randn({8, 4, 1, 1}));
  ar.write("bin_conv_tr1.weight", torch::randn({4, 4, 2, 2}));
  ar.write("bin_conv_tr1.bias", torch::randn({4}));
  // the de-duplicated names of the char-rec net (char_recognition/model.rs:14-17)
  ar.write("bias", torch::randn({3}));
  ar.write("weight", torch::randn({3, 1, 5, 5}));
  ar.write("bias__2", torch::randn({2}));
  ar.write("weight__3", torch::randn({2, 3}));
  // storage kinds and layouts the reader must cope with
  ar.write("f64", torch::arange(6, torch::kDouble).reshape({2, 3}) / 7.0);
  ar.write("f16", (torch::arange(8, torch::kFloat) / 3.0).to(torch::kHalf));
  ar.write("bf16", (torch::arange(8, torch::kFloat) / 3.0).to(torch::kBFloat16));
  ar.write("transposed_view", torch::arange(12, torch::kFloat).reshape({3, 4}).t());
  ar.write("offset_view", torch::arange(20, torch::kFloat).slice(0, 5, 15).reshape({2, 5}));
  ar.write("scalar", torch::tensor(3.25f));
  ar.save_to(argc > 1 ? argv[1] : "tests/golden/varstore_libtorch.ot");
  return 0;
}
