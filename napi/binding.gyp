{
  # node-gyp build of the nodencl-compatible addon over libphaneron_b200.so (napi/phaneron_napi.cc).
  # PHANERON_B200 = the checkout of this repository (include/ and phaneron_b200/libphaneron_b200.so, built by
  # `python -m phaneron_b200.build`).
  "targets": [
    {
      "target_name": "phaneron_b200",
      "sources": ["phaneron_napi.cc"],
      "include_dirs": ["<!@(node -p \"require('node-addon-api').include\")", "<!(echo ${PHANERON_B200:-..})/include"],
      "dependencies": ["<!(node -p \"require('node-addon-api').gyp\")"],
      "cflags_cc": ["-std=c++17", "-fexceptions"],
      "defines": ["NAPI_CPP_EXCEPTIONS"],
      "libraries": ["-L<!(echo ${PHANERON_B200:-..})/phaneron_b200", "-lphaneron_b200", "-Wl,-rpath,<!(echo ${PHANERON_B200:-..})/phaneron_b200"]
    }
  ]
}
