{
  "targets": [{
    "target_name": "ragera_addon",
    "sources": ["ragera_addon.cc"],
    "include_dirs": ["../../include"],
    "libraries": ["-L<(module_root_dir)/../../rag_era_b200", "-lragera", "-Wl,-rpath,<(module_root_dir)/../../rag_era_b200"],
    "cflags_cc": ["-std=c++17", "-O2"],
    "defines": ["NAPI_VERSION=8"]
  }]
}
