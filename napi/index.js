// module entry of the nodencl-compatible addon: the 41 importing files of phaneron keep `from 'nodencl'`
const addon = require('./build/Release/phaneron_b200.node')
module.exports = addon
